"""TEST INFRASTRUCTURE ONLY — pins oracle/head_oracle.py against the reference's own `sample_logistic`
(lib/models/models.py:28-74) and writes tests/golden/head.npz.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_head
lib.models.models imports timm / matplotlib, which this image lacks; they are replaced by inert mock modules — only
`sample_logistic`, a pure torch function, is called.  Stored per case: the reference's fp32 logits, and the reverse rates
the reference's get_reverse_rates (lib/sampling/sampling.py:31-78) makes of them, plus the fp64 evaluation of the same
formulas (the yardstick that says how much fp32 noise the reference itself carries in the 1e-6-guarded log1p).
"""
from __future__ import annotations

import os
import sys
from unittest import mock

import numpy as np
import torch

from . import cases, head_oracle as ho, ref_harness as rh, ctmc_oracle as oc
from .make_golden import _fwd_cfg

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, N, D, fix_logistic, (scale_lo, scale_hi), loss_name, logit_type, t
HEAD_CASES = [
    ("head_tauldr", 4, 48, False, (-3.0, 3.0), "CTElboLambda", None, 0.6),
    ("head_tauldr_fix", 4, 48, True, (-3.0, 3.0), "CTElbo", None, 0.2),
    ("head_revprob", 4, 48, False, (-2.0, 2.0), "CatRM", "reverse_prob", 0.9),
    ("head_revprob_fix", 4, 48, True, (-4.0, 1.0), "SDDMElbo", "reverse_prob", 0.05),
]
FWD = "gauss256"

# whole samplers with the head inside the model: (name, sampler, forward, N, D, loss.name, logit_type, fix_logistic,
# sampler overrides, max_t, seed) — same layout as cases.SAMPLERS with fix_logistic in the stub slot
HEAD_SAMPLERS = [
    ("head_taul_lambda", "TauL", FWD, 8, 12, "CTElboLambda", None, False, dict(num_steps=8, min_t=0.01), 1.0, 301),
    ("head_taul_corr_crm_fix", "TauL", FWD, 8, 12, "CatRM", "reverse_prob", True,
     dict(num_steps=6, min_t=0.01, corrector_entry_time=0.3, num_corrector_steps=2), 1.0, 302),
    ("head_lbjf_ctelbo", "LBJF", FWD, 8, 12, "CTElbo", None, False, dict(num_steps=8, min_t=0.01), 1.0, 303),
    ("head_condtaul", "ConditionalTauLeaping", FWD, 6, 10, "CTElbo", None, False,
     dict(num_steps=6, min_t=0.01, condition_dim=4, reject_multiple_jumps=True), 1.0, 304),
]


def case_inputs(case):
    name, N, D, fix, (lo, hi), loss_name, logit_type, t = case
    seed = sum(map(ord, name))
    mu, ls = ho.head_inputs(N * D, seed, lo, hi)
    S = cases.FORWARD[FWD]["S"]
    g = np.random.Generator(np.random.PCG64(seed))
    # current state near the mode of the head, as in a sampler late in the chain
    centre = np.clip(np.round((mu.numpy() + 1.0) * S / 2.0 - 0.5), 0, S - 1).astype(np.int64)
    x = np.clip(centre + g.integers(-6, 7, N * D), 0, S - 1).reshape(N, D)
    return mu.view(N, D), ls.view(N, D), torch.from_numpy(x), S


def import_reference_head():
    ref = rh.import_reference()
    sys.modules["torchtyping"].patch_typeguard = lambda *a, **k: None
    for _ in range(16):
        try:
            import lib.models.models as mm
            return ref, mm
        except ImportError as e:           # timm, matplotlib, ...: absent here and irrelevant to sample_logistic
            if not e.name:
                raise
            sys.modules[e.name] = mock.MagicMock()
    raise RuntimeError("could not import lib.models.models")


def make_ref_head_model(ref, mm, cfg, S, D, seed, fix):
    """Reference rate mixin + stub network + the REFERENCE's sample_logistic as the output head."""
    mixin = getattr(ref.fm, cases.FORWARD[FWD]["mixin"])

    class M(rh.HeadStubNet, mixin):
        def __init__(self):
            rh.HeadStubNet.__init__(self, S, D, seed)
            mixin.__init__(self, cfg, "cpu")

        def forward(self, x, t):
            mu, ls = self.head_params(x, t)
            N, Dx = mu.shape
            return mm.sample_logistic((mu.view(N, 1, Dx, 1), ls.view(N, 1, Dx, 1)), N, 1, Dx, S, fix, "cpu").reshape(N, Dx, S)

    return M()


def golden_head_samplers(ref, mm, out):
    for case in HEAD_SAMPLERS:
        name, cls, fwd, N, D, loss_name, logit_type, fix, over, max_t, seed = case
        cfg = cases.sampler_cfg(rh.make_cfg, case)
        S = cfg.data.S
        m = make_ref_head_model(ref, mm, cfg, S, D, seed, fix)
        sampler = getattr(ref.ss, cls)(cfg)
        args = ()
        if over.get("condition_dim", 0):
            g = np.random.Generator(np.random.PCG64(seed))
            args = (torch.from_numpy(g.integers(0, S, (N, over["condition_dim"]))),)
        with rh.Injector(ref, seed=seed):
            res = sampler.sample(m, N, *args)
        if not isinstance(res, tuple):
            res = (res,)
        out[f"{name}/x"] = np.asarray(res[0]).astype(np.int64)
        for i, extra in enumerate(res[1:]):
            out[f"{name}/diag{i}"] = np.asarray(extra, dtype=np.float64)
        print(f"  {name}: x mean {out[f'{name}/x'].mean():.3f}")


def main():
    ref, mm = import_reference_head()
    out = {}
    golden_head_samplers(ref, mm, out)
    for case in HEAD_CASES:
        name, N, D, fix, _, loss_name, logit_type, t = case
        mu, ls, x, S = case_inputs(case)
        logits = mm.sample_logistic((mu.view(N, 1, D, 1), ls.view(N, 1, D, 1)), N, 1, D, S, fix, "cpu").reshape(N, D, S)
        mine = ho.truncated_logistic_logits(mu, ls, S, fix)
        assert torch.equal(logits, mine), (name, (logits - mine).abs().max())      # same torch ops -> bitwise
        cfg = _fwd_cfg(FWD)
        cfg["loss"] = rh.Cfg(name=loss_name, logit_type=logit_type)
        cfg["sampler"] = rh.Cfg(eps_ratio=1e-9)
        m = rh.make_ref_model(ref, cases.FORWARD[FWD]["mixin"], cfg, S, D, 0)
        t_ones = t * torch.ones((N,))
        with torch.no_grad():
            rr, ratio = ref.ss.get_reverse_rates(m, logits, x, t_ones, cfg, N, D, S)
            Q = m.transition(t_ones[:1])
            R = m.rate(t_ones[:1])
        l64 = ho.truncated_logistic_logits(mu.double(), ls.double(), S, fix)
        rr64, ratio64 = oc.reverse_rates(l64, x, Q.double(), R.double(), loss_name, logit_type or "reverse_prob")
        out[f"{name}/logits"] = logits.numpy()
        out[f"{name}/rr"] = rr.numpy()
        out[f"{name}/ratio"] = ratio.numpy()
        out[f"{name}/p64"] = torch.softmax(l64, -1).numpy()
        out[f"{name}/rr64"] = rr64.numpy()
        out[f"{name}/ratio64"] = ratio64.numpy()
        p32 = torch.softmax(logits, -1).double()
        print(f"  {name}: reference fp32 vs fp64  p max abs {float((p32 - torch.softmax(l64, -1)).abs().max()):.2e}  "
              f"rr max rel-to-rowmax {float(((rr.double() - rr64).abs().amax(-1) / rr64.abs().amax(-1)).max()):.2e}")
    np.savez_compressed(os.path.join(OUT, "head.npz"), **out)
    print("wrote", os.path.join(OUT, "head.npz"))


if __name__ == "__main__":
    main()
