"""Diagnostic: SDDM-ELBO / CT-ELBO loss terms forward + backward at the C5 shape, tensor-core contractions vs the
CUDA-core kernel (timing and agreement).  Not part of the product."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctdd_b200 import _native as nat, make_config, ops
from ctdd_b200.lib.models import forward_model as fm
B, D, S = int(os.environ.get("B", 64)), 3072, 256
dev = "cuda"
cfg = make_config(data=dict(S=S), model=dict(rate_sigma=6.0, Q_sigma=512.0, time_exp=100.0, time_base=3.0), device=dev)
model = fm.GaussianTargetRate(cfg, dev)
g = torch.Generator(device=dev).manual_seed(1)
ts = torch.rand(B, device=dev, generator=g) * 0.97 + 0.02
Q, QT = model._build_qt0(model._transition_delta(ts), inverse=True, want_transpose=True)
beta = model._rate_scalar(ts).float().contiguous()
Rb, _ = model.base_rate_tables(Q.device)
x0 = torch.randint(0, S, (B, D), device=dev, generator=g, dtype=torch.int32)
xt, xtil = ops.noise_xt(Q, Rb, beta, x0, 5, 0)
logits = (torch.randn(B, D, S, device=dev, generator=g) - (torch.arange(S, device=dev).view(1, 1, S) - x0.unsqueeze(-1)).float() ** 2 / 128.0)
for kind, name in ((nat.LOSS_SDDM, "SDDM"), (nat.LOSS_CTELBO, "CTELBO")):
    res = {}
    for tc in (False, True):
        ops._LossTerms.use_tc = tc
        lg = logits.clone().requires_grad_(True)
        kw = dict(Q=Q, QT=QT, Rb=Rb, beta=beta, x0=x0, eps=1e-9)
        def run():
            if kind == nat.LOSS_SDDM:
                o = ops.loss_terms(lg, kind, xt=xtil, **kw)
            else:
                o = ops.loss_terms(lg, kind, xt=xtil, x_tilde=xtil, **kw)
            loss = (o[0] - o[1] / o[2] + 0.01 * o[4]).mean()
            lg.grad = None
            loss.backward()
            return o, loss
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        times = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o, loss = run()
            e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        res[tc] = (loss.item(), [t.detach().clone() for t in o], lg.grad.clone(), times[len(times) // 2])
    a, b = res[False], res[True]
    gerr = ((a[2] - b[2]).abs().max() / a[2].abs().max()).item()
    oerr = max(((x - y).abs().max() / x.abs().max().clamp_min(1e-30)).item() for x, y in zip(a[1], b[1]))
    print(f"{name} B={B}: loss {a[0]:.6g} vs {b[0]:.6g}; terms rel err {oerr:.2e}; grad rel err {gerr:.2e}; "
          f"fwd+bwd ms (median of 8) cuda-core {a[3]:.3f}  tensor-core {b[3]:.3f}")
