"""CPU: oracle/head_oracle.py against the fixture the reference's own sample_logistic produced (tests/golden/head.npz)."""
import numpy as np
import torch

from oracle import head_oracle as ho, ctmc_oracle as oc
from oracle.make_golden_head import HEAD_CASES, HEAD_SAMPLERS, case_inputs, FWD
from helpers import oracle_forward
import pytest


def test_head_oracle_matches_reference_fixture(golden):
    """Same torch ops as the reference -> probabilities agree to fp32 rounding on any host (the logits themselves carry
    the reference's cancellation noise in the 1e-6-guarded tails, so they are compared after the softmax)."""
    g = golden["head"]
    for case in HEAD_CASES:
        name, N, D, fix = case[:4]
        mu, ls, x, S = case_inputs(case)
        logits = ho.truncated_logistic_logits(mu, ls, S, fix)
        ref = torch.from_numpy(g[f"{name}/logits"])
        assert logits.shape == ref.shape == (N, D, S)
        p, pr = torch.softmax(logits, -1), torch.softmax(ref, -1)
        assert float((p - pr).abs().max()) <= 2e-6
        live = ref > -12.0                      # above the 1e-6 guard: the logits themselves are well conditioned
        assert float((logits - ref)[live].abs().max()) <= 2e-3


def test_head_fp64_yardstick_and_rates(golden):
    """fp64 evaluation of the same formulas reproduces the stored yardstick; the reference's fp32 rates sit within 1e-4."""
    g = golden["head"]
    fp = oracle_forward(FWD)
    for case in HEAD_CASES:
        name, N, D, fix, _, loss_name, logit_type, t = case
        mu, ls, x, S = case_inputs(case)
        l64 = ho.truncated_logistic_logits(mu.double(), ls.double(), S, fix)
        np.testing.assert_allclose(torch.softmax(l64, -1).numpy(), g[f"{name}/p64"], rtol=1e-9, atol=1e-300)
        tt = torch.tensor([t], dtype=torch.float64).to(torch.float32)
        rr, ratio = oc.reverse_rates(ho.truncated_logistic_logits(mu, ls, S, fix), x, fp.transition(tt), fp.rate(tt),
                                     loss_name, logit_type or "reverse_prob")
        ref = g[f"{name}/rr"]
        big = np.abs(ref) > 1e-30
        rel = np.abs(rr.numpy() - ref)[big] / np.abs(ref)[big]
        assert rel.max() <= 1e-4
        r64 = g[f"{name}/rr64"]
        rel64 = np.abs(ref - r64)[big] / np.abs(r64)[big]
        assert rel64.max() <= 1e-4


def test_head_identity_used_by_the_kernels():
    """exp(logits_1) = u_{s+1} (kappa v_s + 1e-6) and exp(min(logits_1, logits_2)) = kappa u_{s+1} v_s + 1e-6 min(u_{s+1}, v_s)
    (the closed form the CUDA producers evaluate) — checked in fp64 against the reference formulas."""
    S = 256
    mu, ls = ho.head_inputs(64, 5, -4.0, 4.0)
    mu, ls = mu.double(), ls.double()
    inv = torch.exp(2.0 - ls).unsqueeze(-1)
    edges = torch.linspace(-1.0, 1.0, S + 1, dtype=torch.float64)
    z = (edges - mu.unsqueeze(-1)) * inv
    u, v = torch.sigmoid(z), torch.sigmoid(-z)
    kap = -torch.expm1(-inv * 2.0 / S)
    p1 = u[:, 1:] * (kap * v[:, :-1] + 1e-6)
    pf = kap * u[:, 1:] * v[:, :-1] + 1e-6 * torch.minimum(u[:, 1:], v[:, :-1])
    for fix, mine in ((False, p1), (True, pf)):
        ref = torch.exp(ho.truncated_logistic_logits(mu, ls, S, fix))
        rel = ((mine - ref).abs() / ref)
        assert float(rel.max()) <= 1e-9, (fix, float(rel.max()))


@pytest.mark.parametrize("case", HEAD_SAMPLERS, ids=[c[0] for c in HEAD_SAMPLERS])
def test_oracle_samplers_with_head_match_reference(golden, case):
    """Whole samplers whose model ends in the truncated-logistic head: the oracle (oracle head + oracle sampler) lands on
    the states the reference (its sample_logistic + its sampler) produced under the same injected uniforms."""
    from helpers import run_oracle_sampler as _run_oracle_sampler
    with torch.no_grad():
        res = _run_oracle_sampler(case, head=True)
    want = golden["head"][f"{case[0]}/x"]
    assert float((np.asarray(res[0]) != want).mean()) <= 0.011      # same torch ops; allows one tie in ~90 states
