"""Shared test helpers: build product models (StubNet x ctdd_b200 rate mixin) and oracle forward processes."""
import numpy as np
import torch

from oracle import cases, ctmc_oracle as oc, ref_harness as rh


def oracle_forward(fwd_name):
    f = cases.FORWARD[fwd_name]
    kw = dict(f["model"])
    return oc.ForwardProcess(f["kind"], f["S"], **kw)


def product_model(fwd_name, cfg, D, seed, scale=0.5, width=None, device="cuda"):
    from ctdd_b200.lib.models import forward_model as fm
    f = cases.FORWARD[fwd_name]
    mixin = getattr(fm, f["mixin"])
    S = f["S"]

    class M(rh.StubNet, mixin):
        def __init__(self):
            rh.StubNet.__init__(self, S, D, seed, scale, width)
            mixin.__init__(self, cfg, device)

        def forward(self, x, t):
            return self.net(x, t)

    m = M().to(device)
    m.device = device
    return m


def fwd_cfg(fwd_name, make_cfg, device="cuda"):
    f = cases.FORWARD[fwd_name]
    model = dict(f["model"])
    model.setdefault("Q_sigma", 20.0)
    return make_cfg(data=dict(S=f["S"]), model=model, device=device)


def mismatch_fraction(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a != b).mean())


RATE_TOL = 1e-4     # north_star: reverse rates within 1e-4 relative; a decision may differ only inside that band


def assert_only_ties(got, want, margin, frac_tol=1e-3, rate_tol=RATE_TOL, what=""):
    """Integer states equal to the oracle's except threshold ties: every mismatching row must have a deciding uniform
    within `rate_tol` (relative rate change) of a threshold of the uniform -> sample map (oracle/rng.py *_margin), and
    at most `frac_tol` of the rows (at least 2) may be such ties."""
    got, want = np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)
    assert got.shape == want.shape, (got.shape, want.shape)
    bad = np.flatnonzero(got != want)
    assert bad.size <= max(2, frac_tol * got.size), (what, bad.size, got.size)
    if bad.size:
        m = np.asarray(margin).reshape(-1)[bad]
        assert np.all(m <= rate_tol), (what, "mismatch that is not a threshold tie", bad[m > rate_tol][:8], m[m > rate_tol][:8])


def run_oracle_sampler(case, head=False):
    name, cls, fwd, N, D, loss_name, logit_type, stub, over, max_t, seed = case
    cfg = cases.sampler_cfg(rh.make_cfg, case)
    S, sc = cfg.data.S, cfg.sampler
    fp = oracle_forward(fwd)
    if head:        # stub network emitting (mu, log_scale) + the oracle's truncated-logistic head; `stub` = fix_logistic
        from oracle import head_oracle as ho
        hnet = rh.HeadStubNet(S, D, seed)
        model = lambda x, t: ho.truncated_logistic_logits(*hnet.head_params(x, t), S, stub)
    else:
        net = rh.StubNet(S, D, seed, stub[0], stub[1])
        model = lambda x, t: net.net(x, t)
    lt = logit_type or "reverse_prob"
    common = dict(min_t=sc.min_t, num_steps=sc.num_steps, initial_dist=sc.initial_dist, seed=seed)
    if cls == "TauL":
        return oc.sample_taul(fp, model, N, D, S, max_t=max_t, init_std=cfg.model.Q_sigma, is_ordinal=sc.is_ordinal,
                              loss_name=loss_name, logit_type=lt, corrector_entry_time=sc.corrector_entry_time,
                              num_corrector_steps=sc.num_corrector_steps, **common)
    if cls == "LBJF":
        return oc.sample_lbjf(fp, model, N, D, S, max_t=max_t, init_std=cfg.model.Q_sigma, loss_name=loss_name,
                              logit_type=lt, corrector_entry_time=sc.corrector_entry_time,
                              num_corrector_steps=sc.num_corrector_steps, **common)
    if cls == "MidPointTauL":
        return oc.sample_midpoint(fp, model, N, D, S, max_t=max_t, init_std=cfg.model.Q_sigma, is_ordinal=sc.is_ordinal,
                                  loss_name=loss_name, logit_type=lt, **common)
    if cls == "ExactSampling":
        return oc.sample_exact(fp, model, N, D, S, max_t=max_t, min_t=sc.min_t, num_steps=sc.num_steps,
                               initial_dist=sc.initial_dist, init_std=cfg.model.Q_sigma, seed=seed)
    if cls == "PCTauL":
        return (oc.sample_pctaul(fp, model, N, D, S, corrector_entry_time=sc.corrector_entry_time,
                                 num_corrector_steps=sc.num_corrector_steps,
                                 corrector_step_size_multiplier=sc.corrector_step_size_multiplier, **common),)
    g = np.random.Generator(np.random.PCG64(seed))
    conditioner = torch.from_numpy(g.integers(0, S, (N, sc.condition_dim)))
    if cls == "ConditionalTauLeaping":
        return (oc.sample_conditional_taul(fp, model, N, D, S, conditioner, condition_dim=sc.condition_dim,
                                           init_std=cfg.model.Q_sigma, **common),)
    return (oc.sample_conditional_pctaul(fp, model, N, D, S, conditioner, condition_dim=sc.condition_dim,
                                         init_std=cfg.model.Q_sigma, reject=bool(sc.reject_multiple_jumps),
                                         corrector_entry_time=sc.corrector_entry_time,
                                         num_corrector_steps=sc.num_corrector_steps,
                                         corrector_step_size_multiplier=sc.corrector_step_size_multiplier, **common),)
