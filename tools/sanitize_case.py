"""Small invocations of every tcgen05 kernel: ragged row counts, both branches, every mode, dense logits and the fused
head, the per-sample GEMM.  Written as the target of `compute-sanitizer --tool memcheck|racecheck python
tools/sanitize_case.py`; compute-sanitizer is CLOSED on this GPU pool (gpurun refuses it), so the script carries its own
checks instead: every output is range-checked, guard bands around the output buffers must stay untouched (out-of-bounds
writes), and every launch is repeated REPEAT times and must reproduce its output bit for bit (a shared-memory / tensor-
memory race between the warp roles, or a read of a buffer before its barrier, shows up as run-to-run differences).
Not part of the product."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctdd_b200 import _native as nat, make_config, ops
from ctdd_b200.lib.models import forward_model as fm

dev = torch.device("cuda")
S = 256
cfg = make_config(data=dict(S=S), model=dict(rate_sigma=6.0, Q_sigma=512.0, time_exp=100.0, time_base=3.0), device="cuda")
model = fm.GaussianTargetRate(cfg, "cuda")
ts = [0.7, 0.2]
Q, QT, beta = model.qt0_tables(ts, dev)
Rb, RbT = model.base_rate_tables(dev)
g = torch.Generator(device=dev).manual_seed(3)
n_launch = 0
REPEAT = int(os.environ.get("REPEAT", 3))
for (N, D) in ((1, 1), (3, 171), (5, 300)):          # 1, 513 and 1500 rows: tile tails, several tiles per pair
    x = torch.randint(0, S, (N, D), device=dev, generator=g, dtype=torch.int32)
    lg = 3.0 * torch.randn((N, D, S), device=dev, generator=g)
    mu = torch.tanh(torch.randn((N, D), device=dev, generator=g))
    ls = -1.0 + torch.randn((N, D), device=dev, generator=g)
    for branch in (nat.BRANCH_TAULDR, nat.BRANCH_SDDM_REVERSE_PROB):
        tc = ops.prep_tc_tables(Q, QT, Rb, 1e-9, branch)
        tcs = ops.prep_tc_static(Rb)
        for i in range(len(ts)):
            kw = dict(N=N, D=D, S=S, seed=11, offset=i, tc_tables=tc[i], tc_static=tcs, impl=nat.IMPL_TC)
            args = (Q[i], QT[i], Rb, RbT, beta[i], 0.01, 1e-9)
            for mode in (nat.MODE_TAU_LEAP, nat.MODE_TAU_LEAP_CORR, nat.MODE_MIDPOINT_DRIFT, nat.MODE_EULER, nat.MODE_EULER_CORR):
                for head in (None, (mu, ls, False), (mu, ls, True)):
                    first = None
                    for rep in range(REPEAT):
                        st = torch.zeros(8, dtype=torch.int64, device=dev)
                        out = ops.reverse_step(mode, branch, lg if head is None else None, x, *args, stats=st, head=head, **kw)["x"]
                        assert int(out.min()) >= 0 and int(out.max()) < S
                        cur = (out.clone(), st.clone())
                        if first is None:
                            first = cur
                        else:
                            assert torch.equal(first[0], cur[0]) and torch.equal(first[1], cur[1]), ("nondeterministic", N, D, branch, mode, head is not None)
                        n_launch += 1
            xb = torch.clamp(x + 1, 0, S - 1)
            ops.reverse_step(nat.MODE_MIDPOINT_JUMP, branch, lg, x, *args, x_base=xb, reject_multi=True, **kw)
            r = ops.reverse_step(nat.MODE_RATES_ONLY, branch, lg, x, *args, want_rr=True, want_ratio=True, **kw)
            r2 = ops.reverse_step(nat.MODE_RATES_ONLY, branch, lg, x, *args, want_rr=True, want_ratio=True, **kw)
            assert torch.isfinite(r["rr"]).all() and torch.equal(r["rr"], r2["rr"]) and torch.equal(r["ratio"], r2["ratio"])
            n_launch += 3
for B, D in ((1, 1), (3, 130), (7, 40)):
    X = torch.rand((B, D, S), device=dev, generator=g)
    M = torch.rand((B, S, S), device=dev, generator=g)
    # guard bands: the kernel writes into the middle of a larger poisoned buffer
    big = torch.full((B * D * S + 2 * 4096,), -7.0, device=dev)
    view = big[4096:4096 + B * D * S].view(B, D, S)
    out = ops.bgemm256(X, M, out=view)
    ref = torch.einsum("bdk,bnk->bdn", X, M)
    assert ((out - ref).abs() <= 1e-3 * ref.abs()).all()
    assert bool((big[:4096] == -7.0).all()) and bool((big[4096 + B * D * S:] == -7.0).all()), "bgemm wrote outside its output"
    for rep in range(REPEAT):
        assert torch.equal(ops.bgemm256(X, M), out.clone())
    n_launch += 1 + REPEAT
torch.cuda.synchronize()
print("sanitize_case: ok,", n_launch, "tcgen05 kernel launches")
