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
