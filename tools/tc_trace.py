"""Diagnostic (not part of the product): per-tile timeline of the tcgen05 reverse-step kernel's warp roles.
Build with  CTDD_TRACE=1 python continuous-time-diffusion-models-for-discrete-data_b200/build.py --force  first."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ctdd_b200 import _native as nat, make_config, ops
from ctdd_b200.lib.models import forward_model as fm

w = bench.WORKLOADS["C4"]
B = int(os.environ.get("B", 1024)); D, S = w["D"], w["S"]
dev = torch.device("cuda")
cfg = make_config(data=dict(S=S), model=dict(w["model"]), device="cuda")
model = fm.GaussianTargetRate(cfg, "cuda")
ts = [float(os.environ.get("T", "0.1"))]
Q, QT, beta = model.qt0_tables(ts, dev)
Rb, RbT = model.base_rate_tables(dev)
branch = nat.BRANCH_TAULDR
tc = ops.prep_tc_tables(Q, QT, Rb, 1e-9, branch); tcs = ops.prep_tc_static(Rb)
lg, x0 = bench.synth_logits(B, D, S, 1, dev)
x = torch.clamp(x0 + torch.randint(-3, 4, x0.shape, device=dev), 0, S - 1).to(torch.int32)
for rep in range(2):
    ops.reverse_step(nat.MODE_TAU_LEAP, branch, lg, x, Q[0], QT[0], Rb, RbT, beta[0], 0.99 / 1000, 1e-9, N=B, D=D, S=S,
                     seed=1, offset=0, tc_tables=tc[0], tc_static=tcs)
torch.cuda.synchronize()
tr = np.zeros((2, 4, 1024, 8), dtype=np.int64)
L = nat.lib()
L.ctdd_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
assert L.ctdd_debug_trace_read(tr.ctypes.data, tr.nbytes) == 0
np.save("gpurun_out/tc_trace.npy", tr)
t0 = int(os.environ.get("TILE0", 200)); n = int(os.environ.get("NT", 6))
for cta in range(2):
    base = tr[cta, 1 if cta == 0 else 0, t0, 0]
    print(f"--- CTA {cta}: cycles relative to tile {t0}")
    names = {0: "prod ", 1: "mma  ", 2: "epi_L", 3: "epi_R"}
    for i in range(t0, t0 + n):
        for role in range(4):
            ev = tr[cta, role, i]
            print(f"tile {i} {names[role]}", " ".join(f"{int(v - base):7d}" if v else "      -" for v in ev))
    d = np.diff(tr[cta, 0, 100:600, 0])
    print("producer tile period: mean", d.mean(), "min", d.min(), "max", d.max())
