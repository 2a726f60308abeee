"""Diagnostic: run tools/tc_probe.py against an alternative build of the library (tools/exp_libs/*.so, made with
CTDD_DEFINES=... python build.py --force) to isolate pipeline stages."""
import os, sys, runpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctdd_b200 import _native as nat
nat.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [sys.argv[0]]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tc_probe.py"), run_name="__main__")
