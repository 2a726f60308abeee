"""Diagnostic: forward noising x_t ~ q_t|0(.|x0) (and the x~ proposal) at the C4/C5 shape, per batch size.  Not part of the product."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ctdd_b200 import make_config, ops
from ctdd_b200.lib.models import forward_model as fm

dev = torch.device("cuda")
w = bench.WORKLOADS["C4"]
S, D = w["S"], w["D"]
model = fm.GaussianTargetRate(make_config(data=dict(S=S), model=dict(w["model"]), device="cuda"), "cuda")
Rb, _ = model.base_rate_tables(dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for B in (64, 128, 512, 2048):
    ts = torch.rand(B, device=dev) * 0.98 + 0.01
    Q, QT = model._build_qt0(model._transition_delta(ts), inverse=True, want_transpose=True)
    beta = model._rate_scalar(ts).float().contiguous()
    x0 = torch.randint(0, S, (B, D), device=dev, dtype=torch.int32)
    t1 = timed(lambda: ops.noise_xt(Q, Rb, beta, x0, 1, 0, want_tilde=False))
    t2 = timed(lambda: ops.noise_xt(Q, Rb, beta, x0, 1, 0))
    xt, _ = ops.noise_xt(Q, Rb, beta, x0, 1, 0)
    print(f"B={B:5d} D={D} S={S}: x_t {t1*1e3:8.1f} us  (Q rows {B*S*S*4/t1/1e6:7.1f} GB/s)   x_t + x~ {t2*1e3:8.1f} us   checksum {int(xt.long().sum())}")
