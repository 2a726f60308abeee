"""CPU: pin the oracle restatement against fixtures produced by the reference itself (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import cases, ctmc_oracle as oc, ref_harness as rh, rng
from oracle.make_golden import rates_inputs
from helpers import oracle_forward, run_oracle_sampler


@pytest.mark.parametrize("name", list(cases.FORWARD))
def test_forward_process_matches_reference(golden, name):
    g = golden["forward"]
    fp = oracle_forward(name)
    t = torch.tensor(cases.FORWARD_TIMES[name], dtype=torch.float32)
    np.testing.assert_array_equal(fp.base_rate.numpy(), g[f"{name}/base_rate"])
    np.testing.assert_allclose(fp.rate(t).numpy(), g[f"{name}/rate"], rtol=0, atol=0)
    # same torch ops on the same machine class: equal up to BLAS blocking noise
    np.testing.assert_allclose(fp.transition(t).numpy(), g[f"{name}/transition"], rtol=0, atol=2e-7)
    if f"{name}/transit_between" in g:
        np.testing.assert_allclose(fp.transit_between(0.5 * t, t).numpy(), g[f"{name}/transit_between"], rtol=0, atol=2e-7)
    if f"{name}/rate_mat" in g:
        y = torch.from_numpy(g[f"{name}/rate_mat_y"])
        np.testing.assert_allclose(fp.rate_mat(y, t).numpy(), g[f"{name}/rate_mat"], rtol=0, atol=0)


def test_transition_properties():
    fp = oracle_forward("gauss256")
    t = torch.tensor([1e-4, 0.3, 1.0])
    Q = fp.transition(t)
    assert torch.all(Q >= 0)
    np.testing.assert_allclose(Q.sum(-1).numpy(), 1.0, atol=2e-5)
    # q_0 ~ identity
    assert torch.all(torch.diagonal(Q[0]) > 0.9)
    # detailed balance of the Gaussian base rate w.r.t. N(S/2, Q_sigma): pi_i R_ij = pi_j R_ji
    S, R = 256, fp.base_rate.double()
    k = torch.arange(1, S + 1, dtype=torch.float64)
    pi = torch.exp(-((k - S / 2) ** 2) / (2 * 512.0 ** 2))
    flow = pi[:, None] * R
    off = ~torch.eye(S, dtype=torch.bool)
    np.testing.assert_allclose(flow[off].numpy(), flow.T[off].numpy(), rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("case", cases.RATES, ids=[c[0] for c in cases.RATES])
def test_reverse_rates_match_reference(golden, case):
    name, fwd, N, D, loss_name, logit_type, stub, t = case
    logits, x, S = rates_inputs(case)
    fp = oracle_forward(fwd)
    tt = t * torch.ones((1,))
    rr, ratio = oc.reverse_rates(logits, x, fp.transition(tt), fp.rate(tt), loss_name, logit_type or "reverse_prob", 1e-9)
    g = golden["rates"]
    np.testing.assert_allclose(rr.numpy(), g[f"{name}/rr"], rtol=2e-5, atol=1e-30)
    np.testing.assert_allclose(ratio.numpy(), g[f"{name}/ratio"], rtol=2e-5, atol=1e-30)


@pytest.mark.parametrize("case", cases.SAMPLERS, ids=[c[0] for c in cases.SAMPLERS])
def test_sampler_matches_reference_with_injected_uniforms(golden, case):
    """Final integer states and diagnostics of the oracle samplers equal the reference's, bit for bit."""
    name = case[0]
    res = run_oracle_sampler(case)
    g = golden["samplers"]
    np.testing.assert_array_equal(np.asarray(res[0]), g[f"{name}/x"])
    for i, extra in enumerate(res[1:]):
        np.testing.assert_allclose(np.asarray(extra, dtype=np.float64), g[f"{name}/diag{i}"], rtol=1e-6, equal_nan=True)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    out = rng.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(w) for w in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = rng.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(w) for w in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = rng.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(w) for w in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_poisson_map_is_poisson():
    """The uniform -> count map reproduces Poisson pmfs (chi-square-free check on a fine uniform grid)."""
    from scipy.stats import poisson
    n = 1 << 20
    v = ((np.arange(n, dtype=np.float64) + 0.5) / n).astype(np.float32)
    for lam in [1e-4, 0.02, 0.7, 3.3, 17.0, 60.0]:
        k = rng.poisson_from_unit(np.full(n, lam, np.float32), v)
        assert abs(k.mean() - lam) < 2e-3 * max(lam, 1e-2) + 4.0 / n * 10
        for kk in range(0, int(lam + 4 * np.sqrt(lam) + 3)):
            assert abs((k == kk).mean() - poisson.pmf(kk, lam)) < 5e-6 + 1e-4 * poisson.pmf(kk, lam)
    # large-rate branch (normal approximation, documented): mean/variance within 1%
    k = rng.poisson_from_unit(np.full(n, 500.0, np.float32), v)
    assert abs(k.mean() - 500.0) < 1.0 and abs(k.var() - 500.0) < 10.0
