"""Diagnostic: per-time-point kernel time of the reverse step at the C4 shape (not part of the product)."""
import sys, os, json
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
ts = [float(v) for v in os.environ.get("TS", "1.0,0.9,0.7,0.5,0.3,0.1,0.05,0.01").split(",")]
Q, QT, beta = model.qt0_tables(ts, dev)
Rb, RbT = model.base_rate_tables(dev)
branch = nat.BRANCH_TAULDR
tc = ops.prep_tc_tables(Q, QT, Rb, 1e-9, branch); tcs = ops.prep_tc_static(Rb)
ws = torch.empty((int(nat.lib().ctdd_step_workspace_bytes(B * D, S, 0)),), dtype=torch.uint8, device=dev)
kind = os.environ.get("LOGITS", "L2")
lg, x0 = bench.synth_logits(B, D, S, 1, dev)
x = torch.clamp(x0 + torch.randint(-3, 4, x0.shape, device=dev), 0, S - 1).to(torch.int32)
h = 0.99 / 1000
for i, t in enumerate(ts):
    if kind == "POST":   # logits of the exact posterior for a flat prior: log q_{t|0}(x_t | x0 = k)
        lg = torch.log(QT[i][x.long()] + 1e-20)
    st = torch.zeros(8, dtype=torch.int64, device=dev)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.reverse_step(nat.MODE_TAU_LEAP, branch, lg, x, Q[i], QT[i], Rb, RbT, beta[i], h, 1e-9, N=B, D=D, S=S,
                               seed=1, offset=i, tc_tables=tc[i], tc_static=tcs, workspace=ws, stats=st if rep == 0 else None)
        e1.record(); torch.cuda.synchronize()
    s = st.cpu().numpy()
    print(f"t={t:5.2f} beta={beta[i]:8.1f} ms={e0.elapsed_time(e1):7.3f} rows_jumped={s[3]/(B*D):.3f} multi={s[4]/(B*D):.3f} changed={s[0]/(B*D):.3f}")
