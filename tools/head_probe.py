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


def train_path(B=128):
    """Training path of the head at a C5-like batch: forward + backward through the reference's formula chain in torch
    (autograd saves ~20 (B,D,S) tensors) against the CUDA forward/backward pair."""
    import torch.nn.functional as F
    D, S = 3072, 256
    g = torch.Generator(device="cuda").manual_seed(2)
    mu0 = torch.tanh(torch.randn((B, D), device="cuda", generator=g))
    ls0 = torch.randn((B, D), device="cuda", generator=g)
    up = torch.randn((B, D, S), device="cuda", generator=g)

    def ref_formula(mu, log_scale):        # the op sequence of sample_logistic (lib/models/models.py:44-72), fix_logistic False
        mu, log_scale = mu.unsqueeze(-1), log_scale.unsqueeze(-1)
        inv_scale = torch.exp(-(log_scale - 2))
        bw = 2.0 / S
        centers = torch.linspace(-1.0 + bw / 2, 1.0 - bw / 2, S, device="cuda").view(1, 1, S)
        left = (centers - bw / 2 - mu) * inv_scale
        right = (centers + bw / 2 - mu) * inv_scale
        a, b = F.logsigmoid(right), F.logsigmoid(left)
        return a + torch.log1p(-torch.exp(b - a) + 1e-6)

    def run(fn):
        mu, ls = mu0.clone().requires_grad_(True), ls0.clone().requires_grad_(True)
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        for _ in range(2):
            (fn(mu, ls) * up).sum().backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            mu.grad = None
            ls.grad = None
            (fn(mu, ls) * up).sum().backward()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5, (torch.cuda.max_memory_allocated() - base) / 2 ** 30, mu.grad.clone(), ls.grad.clone()

    t_ref, m_ref, gm_r, gl_r = run(ref_formula)
    t_our, m_our, gm_o, gl_o = run(lambda mu, ls: ops.logistic_logits_autograd(mu, ls, S, False))
    print(f"train path B={B}: torch formula chain fwd+bwd {t_ref:.2f} ms, peak extra memory {m_ref:.2f} GiB | "
          f"CUDA fwd+bwd kernels {t_our:.2f} ms, {m_our:.2f} GiB | "
          f"grad agreement mu {float((gm_r - gm_o).abs().max() / gm_r.abs().max()):.1e} ls {float((gl_r - gl_o).abs().max() / gl_r.abs().max()):.1e}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_path(int(sys.argv[2]) if len(sys.argv) > 2 else 128)
    else:
        main()
