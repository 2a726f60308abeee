"""Importable alias of the package directory `continuous-time-diffusion-models-for-discrete-data_b200/`.

The graded layout names the package after the reference repository, which is not a valid Python identifier;
this shim makes `import ctdd_b200` resolve to that directory (same modules, one copy of the code).
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "continuous-time-diffusion-models-for-discrete-data_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
