"""Diagnostic: forward + backward time of the fused loss kernels at the C3 / C5 shapes (CUDA events)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctdd_b200 import _native as nat, make_config, ops
from ctdd_b200.lib.models import forward_model as fm

dev = torch.device("cuda")
S = 256
cfg = make_config(data=dict(S=S), model=dict(rate_sigma=6.0, Q_sigma=512.0, time_exp=100.0, time_base=3.0), device="cuda")
model = fm.GaussianTargetRate(cfg, "cuda")
Rb, _ = model.base_rate_tables(dev)
for (name, kind, B, D) in (("C3 CatRMNLL rm", nat.LOSS_CRM, 64, 784), ("C5 SDDMElbo", nat.LOSS_SDDM, 64, 3072),
                           ("C5 CTElbo", nat.LOSS_CTELBO, 64, 3072), ("C5 SDDMElbo B=256", nat.LOSS_SDDM, 256, 3072)):
    ts = torch.rand(B, device=dev) * 0.98 + 0.01
    Q, QT = model._build_qt0(model._transition_delta(ts), inverse=True, want_transpose=True)
    beta = model._rate_scalar(ts).float().contiguous()
    x0 = torch.randint(0, S, (B, D), device=dev, dtype=torch.int32)
    xt, xtil = ops.noise_xt(Q, Rb, beta, x0, 1, 0)
    logits = (torch.randn(B, D, S, device=dev) - (torch.arange(S, device=dev).view(1, 1, S) - x0.unsqueeze(-1)) ** 2 / 128.0).requires_grad_(True)
    def run():
        outs = ops.loss_terms(logits, kind, Q=Q, QT=QT, Rb=Rb, beta=beta, x0=x0, xt=xtil if kind != nat.LOSS_CRM else xt,
                              x_tilde=xtil if kind == nat.LOSS_CTELBO else None, eps=1e-9)
        loss = sum(o.sum() for o in outs)
        loss.backward()
        logits.grad = None
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 6.0 * B * D * S * S
    print(f"{name:22s} B={B:4d} D={D:5d}: fwd+bwd {ms:8.3f} ms  {fl / ms / 1e9:8.2f} TFLOP/s (6*B*D*S^2)  "
          f"{8.0 * B * D * S / ms / 1e6:7.1f} GB/s of logits+grad")
