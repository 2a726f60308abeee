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
