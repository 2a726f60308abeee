"""GPU: the CUDA path (through the C ABI) against the oracle and the reference-generated fixtures."""
import numpy as np
import pytest
import torch

from oracle import cases, ctmc_oracle as oc, ref_harness as rh, rng
from oracle.make_golden import rates_inputs
from helpers import oracle_forward, product_model, fwd_cfg, mismatch_fraction, assert_only_ties, run_oracle_sampler

pytestmark = pytest.mark.gpu

CLAMP = 1e-8


def _nat():
    from ctdd_b200 import _native as nat
    return nat


def _tables(fp, t, dev="cuda"):
    tt = torch.tensor([t], dtype=torch.float64).to(torch.float32)
    Q = fp.transition(tt)[0]
    return dict(Q=Q.to(dev).contiguous(), QT=Q.t().contiguous().to(dev), Rb=fp.base_rate.to(dev).contiguous(),
                RbT=fp.base_rate.t().contiguous().to(dev), beta=float(fp.beta(tt)[0]))


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.FORWARD))
def test_qt0_builder_matches_reference(golden, name):
    """|dQ| <= 5e-7 abs and identical zero pattern outside the clamp band; rows sum to 1 (SURVEY §8c plan (1))."""
    from ctdd_b200 import make_config
    g = golden["forward"]
    cfg = fwd_cfg(name, make_config)
    from ctdd_b200.lib.models import forward_model as fm
    m = getattr(fm, cases.FORWARD[name]["mixin"])(cfg, "cuda")
    t = torch.tensor(cases.FORWARD_TIMES[name], dtype=torch.float32, device="cuda")
    Q = m.transition(t).cpu().numpy()
    ref = g[f"{name}/transition"]
    assert np.abs(Q - ref).max() <= 5e-7
    band = np.abs(ref - CLAMP) < 5e-7
    assert np.array_equal((Q == 0) | band, (ref == 0) | band)
    if name != "uni2":
        np.testing.assert_allclose(Q.sum(-1), 1.0, atol=1e-5)
    np.testing.assert_allclose(m.rate(t).cpu().numpy(), g[f"{name}/rate"], rtol=2e-6, atol=0)
    if f"{name}/transit_between" in g:
        tb = m.transit_between(0.5 * t, t).cpu().numpy()
        assert np.abs(tb - g[f"{name}/transit_between"]).max() <= 2e-6
    if f"{name}/rate_mat" in g:
        y = torch.from_numpy(g[f"{name}/rate_mat_y"]).cuda()
        np.testing.assert_allclose(m.rate_mat(y, t).cpu().numpy(), g[f"{name}/rate_mat"], rtol=2e-6, atol=0)


def test_qt0_tables_transpose_and_cache():
    from ctdd_b200 import make_config
    from ctdd_b200.lib.models import forward_model as fm
    m = fm.GaussianTargetRate(fwd_cfg("gauss32", make_config), "cuda")
    Q, QT, beta = m.qt0_tables([1.0, 0.5, 0.01], "cuda")
    assert torch.equal(Q.transpose(1, 2).contiguous(), QT)
    assert m.qt0_tables([1.0, 0.5, 0.01], "cuda")[0] is Q
    fp = oracle_forward("gauss32")
    np.testing.assert_allclose(beta, fp.beta(torch.tensor([1.0, 0.5, 0.01])).numpy(), rtol=1e-7)


# ---------------------------------------------------------------------------------------------------------------
def _impls():
    nat = _nat()
    return [nat.IMPL_SIMT, nat.IMPL_AUTO]


def _tc(tb, branch, S, impl):
    """tcgen05-path tables for S == 256; IMPL_AUTO is upgraded to IMPL_TC there so a silent SIMT fallback cannot pass."""
    from ctdd_b200 import ops
    nat = _nat()
    if S != 256 or impl == nat.IMPL_SIMT or branch not in (nat.BRANCH_TAULDR, nat.BRANCH_SDDM_REVERSE_PROB):
        return impl, None, None
    tabs = ops.prep_tc_tables(tb["Q"][None], tb["QT"][None], tb["Rb"], 1e-9, branch)
    return nat.IMPL_TC, tabs[0], ops.prep_tc_static(tb["Rb"])


@pytest.mark.parametrize("impl_i", [0, 1])
@pytest.mark.parametrize("case", cases.RATES, ids=[c[0] for c in cases.RATES])
def test_reverse_rates_match_reference(golden, case, impl_i):
    """Reverse rates with the reference's q/R injected: 1e-4 relative, exact zeros preserved (§8c plan (2))."""
    from ctdd_b200 import ops
    nat = _nat()
    name, fwd, N, D, loss_name, logit_type, stub, t = case
    logits, x, S = rates_inputs(case)
    fp = oracle_forward(fwd)
    tb = _tables(fp, t)
    branch = nat.branch_for(loss_name, logit_type)
    impl, tct, tcs = _tc(tb, branch, S, _impls()[impl_i])
    out = ops.reverse_step(nat.MODE_RATES_ONLY, branch, logits.cuda(), x.to(torch.int32).cuda(), tb["Q"], tb["QT"], tb["Rb"],
                           tb["RbT"], tb["beta"], 0.0, 1e-9, N=N, D=D, S=S, impl=impl,
                           tc_tables=tct, tc_static=tcs, want_rr=True, want_ratio=True)
    g = golden["rates"]
    for key, got in (("rr", out["rr"]), ("ratio", out["ratio"])):
        ref = g[f"{name}/{key}"]
        got = got.cpu().numpy()
        big = np.abs(ref) > 1e-30
        rel = np.abs(got - ref)[big] / np.abs(ref)[big]
        assert rel.max() <= 1e-4, (key, rel.max())
        assert np.all(np.abs(got[~big]) <= 1e-30)


def _random_problem(fwd, N, D, t, seed, width, scale=0.5):
    fp = oracle_forward(fwd)
    S = fp.S
    g = np.random.Generator(np.random.PCG64(seed))
    x0 = g.integers(0, S, (N, D))
    s = np.arange(S)
    logits = scale * 4.0 * g.standard_normal((N, D, S))
    if width is not None:
        logits = logits - (s[None, None, :] - x0[:, :, None]) ** 2 / (2.0 * width ** 2)
    x = np.clip(x0 + g.integers(-2, 3, (N, D)), 0, S - 1)
    return fp, torch.from_numpy(logits.astype(np.float32)), torch.from_numpy(x), S


STEP_CASES = [
    # fwd, N, D, t, h, loss, logit_type, width
    ("gauss256", 6, 40, 0.9, 0.004, "CTElbo", None, 12.0),
    ("gauss256", 6, 40, 0.2, 0.01, "CatRM", "reverse_prob", 12.0),
    ("gauss32", 16, 33, 0.5, 0.02, "NLL", None, 3.0),
    ("gauss32", 16, 33, 0.5, 0.02, "CatRM", "direct", 3.0),
    ("gauss32", 16, 33, 0.5, 0.02, "ScoreElbo", "reverse_logscale", 3.0),
    ("univar3_logsqr", 32, 225, 0.4, 0.05, "CTElbo", None, None),
    ("univar2_sqrtcos", 64, 32, 0.6, 0.05, "CatRMNLL", "reverse_prob", None),
    ("univar5_log", 24, 19, 0.6, 0.05, "CTElbo", None, None),
]


@pytest.mark.parametrize("impl_i", [0, 1])
@pytest.mark.parametrize("sc", STEP_CASES, ids=[f"{c[0]}-{c[5]}-{c[6]}" for c in STEP_CASES])
def test_step_modes_match_oracle(sc, impl_i):
    """Every update mode on identical inputs and injected uniforms: integer states bit-exact modulo threshold ties."""
    from ctdd_b200 import ops
    nat = _nat()
    fwd, N, D, t, h, loss_name, logit_type, width = sc
    fp, logits, x, S = _random_problem(fwd, N, D, t, 11, width)
    tb = _tables(fp, t)
    branch = nat.branch_for(loss_name, logit_type)
    impl, tct, tcs = _tc(tb, branch, S, _impls()[impl_i])
    lt = logit_type or "reverse_prob"
    tt = torch.tensor([t], dtype=torch.float64).to(torch.float32)
    Qo, Ro = fp.transition(tt), fp.rate(tt)
    rr, _ = oc.reverse_rates(logits, x, Qo, Ro, loss_name, lt, 1e-9)
    rz = oc._zero_at(rr, x)
    rz_corr = oc._zero_at(Ro.expand(N, -1, -1)[torch.arange(N).view(N, 1), x.long()] + rz, x)
    xe = x.to(torch.int32).cuda()
    lg = logits.cuda()
    seed = 4242

    def run(mode, offset, reject=False, x_base=None):
        stats = torch.zeros(8, dtype=torch.int64, device="cuda")
        use = impl            # every update mode runs on the tcgen05 path at S=256
        out = ops.reverse_step(mode, branch, lg, xe, tb["Q"], tb["QT"], tb["Rb"], tb["RbT"], tb["beta"], h, 1e-9,
                               N=N, D=D, S=S, impl=use, tc_tables=tct, tc_static=tcs,
                               reject_multi=reject, seed=seed, offset=offset, x_base=x_base, stats=stats)
        return out["x"].cpu().numpy().astype(np.int64), stats.cpu().numpy()

    tol = 1e-3
    m_jump3 = oc.tau_leap_margin(rz, h, 3, seed)
    for reject in (False, True):
        got, st = run(nat.MODE_TAU_LEAP, 3, reject)
        want, ost = oc.tau_leap_update(rz, x, x, h, S, reject, 3, seed)
        assert_only_ties(got, want.numpy(), m_jump3, tol, what="tau_leap")
        assert abs(st[nat.STAT_CHANGED_BASE] - ost["changed_base"]) <= max(2, tol * N * D)
        assert abs(st[nat.STAT_ROWS_MULTI] - ost["rows_multi"]) <= max(2, tol * N * D)
    got, _ = run(nat.MODE_TAU_LEAP_CORR, 5)
    want, _ = oc.tau_leap_update(rz_corr, x, x, h, S, False, 5, seed)
    assert_only_ties(got, want.numpy(), oc.tau_leap_margin(rz_corr, h, 5, seed), tol, what="tau_leap_corr")
    got, _ = run(nat.MODE_MIDPOINT_DRIFT, 0)
    want = oc.midpoint_drift(rz, x, h, S)
    assert_only_ties(got, want.numpy(), oc.midpoint_drift_margin(rz, x, h, S), tol, what="midpoint_drift")
    g = np.random.Generator(np.random.PCG64(5))
    xb = torch.from_numpy(np.clip(x.numpy() + g.integers(-1, 2, x.shape), 0, S - 1))
    got, st = run(nat.MODE_MIDPOINT_JUMP, 7, True, xb.to(torch.int32).cuda())
    want, ost = oc.tau_leap_update(rz, x, xb, h, S, True, 7, seed)
    assert_only_ties(got, want.numpy(), oc.tau_leap_margin(rz, h, 7, seed), tol, what="midpoint_jump")
    assert abs(st[nat.STAT_NONZERO_JUMP] - ost["nonzero_jump"]) <= max(2, tol * N * D)
    got, _ = run(nat.MODE_EULER, 9)
    want, _ = oc.euler_update(rz, x, h, S, 9, seed)
    assert_only_ties(got, want.numpy(), oc.euler_margin(rz, x, h, S, 9, seed), tol, what="euler")
    got, _ = run(nat.MODE_EULER_CORR, 10)
    want, _ = oc.euler_update(rz_corr, x, h, S, 10, seed)
    assert_only_ties(got, want.numpy(), oc.euler_margin(rz_corr, x, h, S, 10, seed), tol, what="euler_corr")


def test_row_offset_sharding_is_invariant():
    """Philox is keyed on the GLOBAL row: two half-batches with row offsets == one full batch (multi-GPU invariance)."""
    from ctdd_b200 import ops
    nat = _nat()
    fp, logits, x, S = _random_problem("gauss32", 16, 24, 0.5, 3, 3.0)
    tb = _tables(fp, 0.5)
    kw = dict(Q=tb["Q"], QT=tb["QT"], Rb=tb["Rb"], RbT=tb["RbT"], beta=tb["beta"], h=0.05, eps=1e-9, S=S, D=24, seed=9, offset=2)
    full = ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, logits.cuda(), x.to(torch.int32).cuda(), N=16, **kw)["x"]
    lo = ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, logits[:8].cuda(), x[:8].to(torch.int32).cuda(), N=8, **kw)["x"]
    hi = ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, logits[8:].cuda(), x[8:].to(torch.int32).cuda(), N=8,
                          row_offset=8 * 24, **kw)["x"]
    assert torch.equal(full, torch.cat([lo, hi]))
    with pytest.raises(RuntimeError):
        ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, logits[8:].cuda(), x[8:].to(torch.int32).cuda(), N=8,
                         row_offset=3, **kw)


@pytest.mark.parametrize("fwd,mode_name", [("univar2_sqrtcos", "tau_leap"), ("univar3_logsqr", "euler"),
                                           ("univar3_logsqr", "tau_leap"), ("univar5_log", "tau_leap")])
def test_small_state_kernels_ragged_tail_and_row_offset(fwd, mode_name):
    """The S <= 8 kernels (the lean S = 2 tau-leap / S = 3 Euler instantiations and the general one): a row count that is
    not a multiple of the 8 rows a thread owns, a non-zero row offset, and the shared count call of 4 consecutive rows
    across the shard boundary - against the oracle on the same uniforms, and full batch == two shards."""
    from ctdd_b200 import ops
    nat = _nat()
    N, D, t, h, seed, off = 13, 21, 0.5, 0.08, 31, 6          # 273 rows = 34 groups of 8 + 1 row
    fp, logits, x, S = _random_problem(fwd, N, D, t, 23, None)
    tb = _tables(fp, t)
    tt = torch.tensor([t], dtype=torch.float64).to(torch.float32)
    rr, _ = oc.reverse_rates(logits, x, fp.transition(tt), fp.rate(tt), "CTElbo", "reverse_prob", 1e-9)
    rz = oc._zero_at(rr, x)
    mode = nat.MODE_EULER if mode_name == "euler" else nat.MODE_TAU_LEAP
    kw = dict(Q=tb["Q"], QT=tb["QT"], Rb=tb["Rb"], RbT=tb["RbT"], beta=tb["beta"], h=h, eps=1e-9, S=S, D=D, seed=seed, offset=off)
    row0 = 8 * D                                               # this batch sits behind 8 samples of another rank
    got = ops.reverse_step(mode, nat.BRANCH_TAULDR, logits.cuda(), x.to(torch.int32).cuda(), N=N, row_offset=row0, **kw)["x"]
    if mode_name == "euler":
        want, _ = oc.euler_update(rz, x, h, S, off, seed, row_offset=row0)
        margin = oc.euler_margin(rz, x, h, S, off, seed, row_offset=row0)
    else:
        want, _ = oc.tau_leap_update(rz, x, x, h, S, False, off, seed, row_offset=row0)
        margin = oc.tau_leap_margin(rz, h, off, seed, row_offset=row0)
    assert_only_ties(got.cpu().numpy().astype(np.int64), want.numpy(), margin, 1e-3, what=f"{fwd} {mode_name} ragged")
    assert (got.cpu().numpy().astype(np.int64) != x.numpy()).mean() > 0.02          # the step does move states
    lo = ops.reverse_step(mode, nat.BRANCH_TAULDR, logits[:8].cuda(), x[:8].to(torch.int32).cuda(), N=8, row_offset=row0, **kw)["x"]
    hi = ops.reverse_step(mode, nat.BRANCH_TAULDR, logits[8:].cuda(), x[8:].to(torch.int32).cuda(), N=N - 8,
                          row_offset=row0 + 8 * D, **kw)["x"]
    assert torch.equal(got, torch.cat([lo, hi]))


def test_initial_samples_and_noising_match_oracle():
    from ctdd_b200 import ops
    from ctdd_b200.lib.sampling import sampling
    for S, dist, std in ((256, "gaussian", 512.0), (3, "uniform", None), (32, "gaussian", 4.0)):
        x = sampling.get_initial_samples(50, 37, "cuda", S, dist, std, seed=77).cpu().numpy()
        want = oc.initial_samples(50, 37, S, dist, std, 77).numpy()
        assert mismatch_fraction(x, want) <= 1e-3
    from oracle import loss_oracle as lo
    # D >= S: the CTA-per-sample kernel (each row of Q[b] scanned once); D < S: the lane-per-row kernel; gauss256 with
    # D = 300: two passes of 128 staged rows
    for fwd, B, D in (("gauss32", 12, 50), ("gauss32", 12, 20), ("gauss256", 5, 300)):
        fp = oracle_forward(fwd)
        S = fp.S
        g = np.random.Generator(np.random.PCG64(1))
        x0 = torch.from_numpy(g.integers(0, S, (B, D)))
        ts = torch.from_numpy(g.uniform(0.01, 1.0, B).astype(np.float32))
        Q = fp.transition(ts)
        beta = fp.beta(ts)
        xt, xtil = ops.noise_xt(Q.cuda(), fp.base_rate.cuda(), beta.cuda(), x0.to(torch.int32).cuda(), seed=5, offset=3)
        wxt, wtil = lo.noise_xt(Q, fp.rate(ts), x0, seed=5, offset=3)
        assert mismatch_fraction(xt.cpu().numpy(), wxt.numpy()) <= 1e-3, (fwd, D)
        assert mismatch_fraction(xtil.cpu().numpy(), wtil.numpy()) <= 5e-3, (fwd, D)


# ---------------------------------------------------------------------------------------------------------------
def _run_product_sampler(case, inject_oracle_q, impl=None):
    from ctdd_b200 import make_config
    from ctdd_b200.lib.sampling import sampling_utils
    import ctdd_b200.lib.sampling.sampling  # noqa: F401
    name, cls, fwd, N, D, loss_name, logit_type, stub, over, max_t, seed = case
    cfg = cases.sampler_cfg(make_config, case)
    cfg.device = "cuda"
    S = cfg.data.S
    m = product_model(fwd, cfg, D, seed, stub[0], stub[1])
    if inject_oracle_q:
        fp = oracle_forward(fwd)

        def tables(ts, device):
            t32 = torch.tensor([float(t) for t in ts], dtype=torch.float64).to(torch.float32)
            Q = fp.transition(t32)
            return Q.to(device).contiguous(), Q.transpose(1, 2).contiguous().to(device), [float(b) for b in fp.beta(t32)]

        m.qt0_tables = tables
        # ExactSampling asks the mixin directly (transition / transit_between of two time vectors)
        m.transition = lambda t: fp.transition(t.detach().cpu().float()).to(t.device)
        m.transit_between = lambda t1, t2: fp.transit_between(t1.detach().cpu().float(), t2.detach().cpu().float()).to(t1.device)
    cfg.sampler.name = cls
    sampler = sampling_utils.get_sampler(cfg)
    sampler.seed = seed
    if impl is not None:
        sampler.impl = impl
    args = ()
    if "condition_dim" in over:
        g = np.random.Generator(np.random.PCG64(seed))
        args = (torch.from_numpy(g.integers(0, S, (N, over["condition_dim"]))),)
    res = sampler.sample(m, N, *args)
    return res if isinstance(res, tuple) else (res,)


@pytest.mark.parametrize("case", cases.SAMPLERS, ids=[c[0] for c in cases.SAMPLERS])
def test_samplers_match_reference_fixtures(golden, case):
    """Whole reverse process vs the reference's own sampler output (injected uniforms, reference q injected):
    final integer states bit-exact except threshold ties (§8c plan (3)); diagnostics agree."""
    name = case[0]
    res = _run_product_sampler(case, inject_oracle_q=True)
    g = golden["samplers"]
    assert res[0].dtype.kind == "i" and res[0].shape == g[f"{name}/x"].shape
    assert mismatch_fraction(res[0], g[f"{name}/x"]) <= 1e-3     # observed 0; a tie early in a run moves later steps
    for i, extra in enumerate(res[1:]):
        np.testing.assert_allclose(np.asarray(extra, dtype=np.float64), g[f"{name}/diag{i}"], atol=0.02, rtol=0.05,
                                   equal_nan=True)


@pytest.mark.parametrize("case", cases.SAMPLERS_S256, ids=[c[0] for c in cases.SAMPLERS_S256])
def test_samplers_at_s256_match_oracle(case):
    """BASELINE config C5's sampler (MidPointTauL: drift + jump evaluation per step) and LBJF at S = 256 on the tcgen05
    path, whole reverse process with injected uniforms against the oracle samplers (the reference's own MidPointTauL
    cannot run for DiscreteCIFAR10; the oracle is pinned to it at S = 2 / 3): final states equal up to 2 threshold ties
    of 384, diagnostics agree."""
    res = _run_product_sampler(case, inject_oracle_q=True)
    want = run_oracle_sampler(case)
    got_x, want_x = np.asarray(res[0]), np.asarray(want[0])
    assert got_x.shape == want_x.shape and got_x.dtype.kind == "i"
    assert int((got_x != want_x).sum()) <= 2, (got_x != want_x).sum()
    for a, b in zip(res[1:], want[1:]):
        np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), atol=0.02, rtol=0.05,
                                   equal_nan=True)


@pytest.mark.parametrize("case", [cases.SAMPLERS[0], cases.SAMPLERS[4], cases.SAMPLERS[6]], ids=lambda c: c[0])
def test_samplers_native_q_close_to_reference(golden, case):
    """Same, with q_{t|0} from the CUDA builder: clamp-band flips may move a few states (SURVEY §7 hard part 3)."""
    name = case[0]
    res = _run_product_sampler(case, inject_oracle_q=False)
    assert mismatch_fraction(res[0], golden["samplers"][f"{name}/x"]) <= 0.03


def test_tc_path_full_tiles_and_tail_match_simt():
    """tcgen05 path vs the CUDA-core path on the same inputs (both through the C ABI): ragged row counts (tile tails),
    several tiles per CTA, both branches; rates within 1e-4, states identical up to threshold ties."""
    from ctdd_b200 import ops
    nat = _nat()
    for (N, D) in ((1, 1), (3, 21), (7, 640), (64, 300), (48, 2048)):   # the last: > RING tiles per CTA pair
        for loss_name, lt in (("CTElbo", None), ("CatRM", "reverse_prob")):
            fp, logits, x, S = _random_problem("gauss256", N, D, 0.35, 17 + N, 12.0)
            tb = _tables(fp, 0.35)
            branch = nat.branch_for(loss_name, lt)
            _, tct, tcs = _tc(tb, branch, S, nat.IMPL_AUTO)
            kw = dict(N=N, D=D, S=S, tc_tables=tct, tc_static=tcs)
            args = (branch, logits.cuda(), x.to(torch.int32).cuda(), tb["Q"], tb["QT"], tb["Rb"], tb["RbT"], tb["beta"])
            r_tc = ops.reverse_step(nat.MODE_RATES_ONLY, *args, 0.0, 1e-9, impl=nat.IMPL_TC, want_rr=True, want_ratio=True, **kw)
            r_si = ops.reverse_step(nat.MODE_RATES_ONLY, *args, 0.0, 1e-9, impl=nat.IMPL_SIMT, want_rr=True, want_ratio=True, **kw)
            for key in ("rr", "ratio"):
                a, b = r_tc[key].cpu().numpy(), r_si[key].cpu().numpy()
                big = np.abs(b) > 1e-30
                assert (np.abs(a - b)[big] / np.abs(b)[big]).max() <= 1e-4
                assert np.all(np.abs(a[~big]) <= 1e-30)
            # tie margins from the CUDA-core path's own fp32 rates (s == x zeroed, corrector: + R_t[x, :])
            xl = x.long()
            onehot = torch.nn.functional.one_hot(xl, S).bool()
            rz = r_si["rr"].cpu().masked_fill(onehot, 0.0)
            Rt = (tb["beta"] * tb["Rb"].cpu())[xl]                       # (N, D, S): row x of R_t
            rz_corr = (rz + Rt).masked_fill(onehot, 0.0)
            for mode, reject in ((nat.MODE_TAU_LEAP, False), (nat.MODE_TAU_LEAP, True), (nat.MODE_TAU_LEAP_CORR, False)):
                st1 = torch.zeros(8, dtype=torch.int64, device="cuda")
                st2 = torch.zeros(8, dtype=torch.int64, device="cuda")
                x_tc = ops.reverse_step(mode, *args, 0.01, 1e-9, impl=nat.IMPL_TC, reject_multi=reject, seed=5, offset=1, stats=st1, **kw)["x"]
                x_si = ops.reverse_step(mode, *args, 0.01, 1e-9, impl=nat.IMPL_SIMT, reject_multi=reject, seed=5, offset=1, stats=st2, **kw)["x"]
                margin = oc.tau_leap_margin(rz_corr if mode == nat.MODE_TAU_LEAP_CORR else rz, 0.01, 1, 5)
                assert_only_ties(x_tc.cpu().numpy(), x_si.cpu().numpy(), margin, 1e-3, what=f"tc vs simt N={N} D={D} mode={mode}")
                assert np.abs(st1.cpu().numpy() - st2.cpu().numpy()).max() <= max(2, 2e-3 * N * D)


def test_tc_path_strided_logits_view_equals_contiguous_copy():
    """The conditional samplers hand the kernel the view logits[:, c:, :] of the network's output (pointer offset + batch
    stride, reference sampling.py:369-426) - on the tensor path those rows are not contiguous across samples, so the loader
    warp copies them row by row.  The result must be bit-equal to the same call on a contiguous copy, in every sampling
    mode, for row counts with tile tails."""
    from ctdd_b200 import ops
    nat = _nat()
    for (N, Dfull, c) in ((3, 29, 8), (6, 200, 72), (40, 340, 40)):
        D = Dfull - c
        fp, logits_full, x_full, S = _random_problem("gauss256", N, Dfull, 0.35, 41 + N, 12.0)
        tb = _tables(fp, 0.35)
        for loss_name, lt in (("CTElbo", None), ("CatRM", "reverse_prob")):
            branch = nat.branch_for(loss_name, lt)
            _, tct, tcs = _tc(tb, branch, S, nat.IMPL_AUTO)
            lg = logits_full.cuda()
            x = x_full[:, c:].contiguous().to(torch.int32).cuda()
            lg_copy = lg[:, c:, :].contiguous()
            common = dict(N=N, D=D, S=S, tc_tables=tct, tc_static=tcs, impl=nat.IMPL_TC, seed=11, offset=3)
            for mode in (nat.MODE_TAU_LEAP, nat.MODE_TAU_LEAP_CORR, nat.MODE_MIDPOINT_DRIFT, nat.MODE_EULER):
                a = ops.reverse_step(mode, branch, lg, x, tb["Q"], tb["QT"], tb["Rb"], tb["RbT"], tb["beta"], 0.01, 1e-9,
                                     logits_offset_elems=c * S, batch_stride=Dfull * S, **common)["x"]
                b = ops.reverse_step(mode, branch, lg_copy, x, tb["Q"], tb["QT"], tb["Rb"], tb["RbT"], tb["beta"], 0.01, 1e-9,
                                     **common)["x"]
                assert torch.equal(a, b), (N, Dfull, c, loss_name, mode)


def test_free_running_histograms_vs_reference_law():
    """SURVEY §8c plan (4) / north_star: where bit-exactness cannot be claimed (threshold ties of the 3xBF16 tensor path,
    the chunked superposition map instead of S independent draws, the Cornish-Fisher quantile above lambda = 64) the full
    reverse process must agree in DISTRIBUTION with the reference's law.  The product TauL (tcgen05 path, Philox) against
    the CPU oracle of TauL.sample driven by real torch.poisson draws (sampling.py:131), same stub network, independent
    randomness, N = 16 384: per-dimension state histograms (16 bins of 16 states) have symmetrised KL < 1e-3 on average
    (sampling noise of two N-samples ~ 15/N ~ 9e-4 at the bound, observed well below) and < 3e-3 in every dimension."""
    from ctdd_b200 import make_config
    from ctdd_b200.lib.sampling import sampling_utils
    import ctdd_b200.lib.sampling.sampling  # noqa: F401
    nat = _nat()
    S, D, N = 256, 8, 16384
    case = ("hist", "TauL", "gauss256", N, D, "CTElbo", None, (0.3, 12.0), dict(num_steps=24, min_t=0.01), 1.0, 7)
    cfg = cases.sampler_cfg(make_config, case)
    cfg.device = "cuda"
    m = product_model("gauss256", cfg, D, 7, 0.3, 12.0)
    sampler = sampling_utils.get_sampler(cfg)
    sampler.seed = 1001
    sampler.impl = nat.IMPL_AUTO
    x_gpu = np.asarray(sampler.sample(m, N)[0])
    # the oracle with the reference's own Poisson law, on the CPU copy of the same network
    fp = oracle_forward("gauss256")
    net = rh.StubNet(S, D, 7, 0.3, 12.0)
    oc.set_poisson_law("torch", seed=2002)
    try:
        x_ref, _ = oc.sample_taul(fp, lambda xx, tt: net.net(xx, tt), N, D, S, max_t=case[9],
                                  min_t=cfg.sampler.min_t, num_steps=cfg.sampler.num_steps,
                                  initial_dist=cfg.sampler.initial_dist, init_std=cfg.model.Q_sigma,
                                  is_ordinal=cfg.sampler.is_ordinal, loss_name="CTElbo", seed=2002)
    finally:
        oc.set_poisson_law("map")
    hists = []
    for x in (x_gpu, np.asarray(x_ref)):
        assert x.shape == (N, D) and x.min() >= 0 and x.max() < S
        h = np.stack([np.bincount(x[:, d] // 16, minlength=16) for d in range(D)]).astype(np.float64)
        hists.append((h + 0.5) / (h + 0.5).sum(axis=1, keepdims=True))
    p, q = hists
    kl = 0.5 * ((p * np.log(p / q)).sum(axis=1) + (q * np.log(q / p)).sum(axis=1))
    assert kl.mean() < 1e-3, kl
    assert kl.max() < 3e-3, kl


def test_full_size_c4_properties():
    """BASELINE.json's full C4 size (B=1024, D=3072, S=256: 3.1 M rows, 3.2 GB of logits) through size-independent
    properties: states stay in range, the statistics counters equal what the output shows, the run is deterministic in
    (seed, offset), batch sharding with row offsets reproduces the single launch bit for bit, and two 4096-row slices
    agree with the CUDA-core path run on just those rows (Philox is keyed on the global row)."""
    from ctdd_b200 import ops
    nat = _nat()
    N, D, S, t, h = 1024, 3072, 256, 0.3, 0.99 / 1000
    fp = oracle_forward("gauss256")
    tb = _tables(fp, t)
    branch = nat.BRANCH_TAULDR
    _, tct, tcs = _tc(tb, branch, S, nat.IMPL_AUTO)
    g = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.randint(0, S, (N, D), generator=g, device="cuda")
    logits = torch.randn((N, D, S), generator=g, device="cuda")
    logits -= (torch.arange(S, device="cuda", dtype=torch.float32).view(1, 1, S) - x0.unsqueeze(-1).float()) ** 2 / 128.0
    x = torch.clamp(x0 + torch.randint(-3, 4, x0.shape, generator=g, device="cuda"), 0, S - 1).to(torch.int32)
    common = dict(D=D, S=S, tc_tables=tct, tc_static=tcs, seed=99, offset=4)
    args = (branch,)
    tabs = (tb["Q"], tb["QT"], tb["Rb"], tb["RbT"], tb["beta"], h, 1e-9)

    def run(lg, xs, n, impl, row_offset=0, seed=99, stats=None):
        kw = dict(common)
        kw["seed"] = seed
        if impl == nat.IMPL_SIMT:
            kw["tc_tables"] = kw["tc_static"] = None
        return ops.reverse_step(nat.MODE_TAU_LEAP, branch, lg, xs, *tabs, N=n, impl=impl, row_offset=row_offset, stats=stats, **kw)["x"]

    st = torch.zeros(8, dtype=torch.int64, device="cuda")
    full = run(logits, x, N, nat.IMPL_TC, stats=st)
    assert full.shape == (N, D) and int(full.min()) >= 0 and int(full.max()) < S
    changed = int((full != x).sum())
    assert int(st[nat.STAT_CHANGED_BASE]) == changed == int(st[nat.STAT_CHANGED_EVAL])
    assert 0.2 * N * D < int(st[nat.STAT_ROWS_JUMPED]) < 0.8 * N * D and int(st[nat.STAT_ROWS_MULTI]) <= int(st[nat.STAT_ROWS_JUMPED])
    assert changed <= int(st[nat.STAT_ROWS_JUMPED])
    assert torch.equal(full, run(logits, x, N, nat.IMPL_TC))                       # deterministic
    assert not torch.equal(full, run(logits, x, N, nat.IMPL_TC, seed=100))         # and a function of the seed
    half = N // 2                                                                  # sharding invariance at full size
    lo = run(logits[:half], x[:half], half, nat.IMPL_TC)
    hi = run(logits[half:], x[half:], half, nat.IMPL_TC, row_offset=half * D)
    assert torch.equal(full, torch.cat([lo, hi]))
    for n0 in (0, 771):                                                            # slices vs the CUDA-core path
        sl = slice(n0, n0 + 2)                                                     # 2 samples = 6144 rows
        lg_s, x_s = logits[sl].contiguous(), x[sl].contiguous()
        ref = run(lg_s, x_s, 2, nat.IMPL_SIMT, row_offset=n0 * D)
        rr = ops.reverse_step(nat.MODE_RATES_ONLY, branch, lg_s, x_s, *tabs, N=2, D=D, S=S, impl=nat.IMPL_SIMT, want_rr=True)["rr"]
        rz = rr.cpu().masked_fill(torch.nn.functional.one_hot(x_s.long().cpu(), S).bool(), 0.0)
        margin = oc.tau_leap_margin(rz, h, 4, 99, row_offset=n0 * D)
        assert_only_ties(full[sl].cpu().numpy(), ref.cpu().numpy(), margin, 1e-3, what=f"C4 slice {n0}")
