"""Diagnostic: time the tcgen05 reverse step at the C4 shape with dense logits vs the fused truncated-logistic head,
and the standalone head kernel.  python tools/head_probe.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctdd_b200  # noqa: E402
from ctdd_b200 import ops, _native as nat, make_config  # noqa: E402
from ctdd_b200.lib.models import forward_model as fm  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    D, S = 3072, 256
    cfg = make_config(data=dict(S=S), model=dict(rate_sigma=6.0, Q_sigma=512.0, time_exp=100.0, time_base=3.0), device="cuda")
    m = fm.GaussianTargetRate(cfg, "cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randint(0, S, (B, D), device="cuda", generator=g)
    out2 = torch.randn((B, 2 * D), device="cuda", generator=g)
    stat = ops.prep_tc_static(m.base_rate)
    for t in (1.0, 0.5, 0.1, 0.01):
        Q, QT, beta = m.qt0_tables([t], "cuda")
        tabs = ops.prep_tc_tables(Q, QT, m.base_rate, 1e-9, nat.BRANCH_TAULDR)
        # denoiser-like head: mean near the clean pixel, scale growing with the noise level
        mu = torch.tanh((x0.float() + 0.5) / 128.0 - 1.0 + 0.05 * out2[:, :D])
        ls = (-1.5 + 2.5 * t) + 0.3 * out2[:, D:]
        both = torch.cat([mu, ls], 1).contiguous()
        mu_v, ls_v = torch.chunk(both, 2, dim=1)
        x = torch.clamp(x0 + torch.randint(-8, 9, (B, D), device="cuda", generator=g), 0, S - 1).to(torch.int32)
        h = 0.001
        kw = dict(N=B, D=D, S=S, impl=nat.IMPL_TC, tc_tables=tabs[0], tc_static=stat, seed=3)
        args = (Q[0], QT[0], m.base_rate, m.base_rate.t().contiguous(), beta[0], h, 1e-9)

        def timeit(fn, n=5):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        t_head = timeit(lambda: ops.logistic_logits(mu_v, ls_v, S, False))
        logits = ops.logistic_logits(mu_v, ls_v, S, False)
        t_dense = timeit(lambda: ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, logits, x, *args, **kw))
        t_fused = timeit(lambda: ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, None, x, *args, head=(mu_v, ls_v, False), **kw))
        xa = ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, logits, x, *args, **kw)["x"]
        xb = ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, None, x, *args, head=(mu_v, ls_v, False), **kw)["x"]
        diff = float((xa != xb).float().mean())
        del logits
        print(f"t={t:5.2f} head kernel {t_head:6.3f} ms | dense step {t_dense:6.3f} ms | head+dense {t_head + t_dense:6.3f} ms | "
              f"fused {t_fused:6.3f} ms | states differing fused vs dense {diff:.2e} | changed {float((xa != x).float().mean()):.3f}")


if __name__ == "__main__":
    main()
