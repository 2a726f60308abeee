"""CPU: the uniform -> jump-count map used by every tau-leap kernel and by the injected reference (oracle/rng.py
poisson_rows) draws S INDEPENDENT Poisson(lam_s) counts per row (superposition: total ~ Poisson(sum lam), picks ~
Categorical(lam / sum)).  Checked here: marginal pmfs, means, pairwise independence, determinism, sharding."""
import numpy as np
import pytest
from scipy.stats import poisson

from oracle import rng


@pytest.mark.parametrize("extra_states", [0, 4], ids=["S8_shared_count_call", "S12_per_row_call"])
def test_poisson_rows_marginals_and_independence(extra_states):
    """Both maps of the total count's uniform: S <= 8 (word row & 3 of a call shared by 4 consecutive rows) and S > 8
    (word 0 of the row's own call)."""
    lam_row = np.array([0.0, 0.02, 0.7, 0.0, 1.9, 0.3, 0.004, 2.5] + [0.0] * extra_states, dtype=np.float32)
    assert (len(lam_row) <= rng.JUMP_SHARED_MAX_S) == (extra_states == 0)
    n = 400_000
    lam = np.tile(lam_row, (n, 1))
    k, K = rng.poisson_rows(lam, 0, 7, 0xC7DD)
    assert k.shape == lam.shape and np.array_equal(k.sum(1), K)
    assert np.all(k[:, lam_row == 0] == 0)
    se = np.sqrt(lam_row / n) + 1e-9
    assert np.all(np.abs(k.mean(0) - lam_row) < 5 * se)
    for s, l in enumerate(lam_row):
        if l == 0:
            continue
        for kk in range(0, 8):
            p = poisson.pmf(kk, l)
            assert abs((k[:, s] == kk).mean() - p) < 5 * np.sqrt(p * (1 - p) / n) + 2e-5
    # independence: covariances vanish (a multinomial split WITHOUT the Poisson total would give -n p_i p_j)
    c = np.cov(k[:, lam_row > 0].T.astype(np.float64))
    off = c - np.diag(np.diag(c))
    assert np.abs(off).max() < 0.012
    np.testing.assert_allclose(np.diag(c), lam_row[lam_row > 0], rtol=0.03, atol=2e-4)
    # joint check on one pair: P(k_2 = 0, k_4 = 0) = exp(-(lam_2 + lam_4))
    both0 = ((k[:, 2] == 0) & (k[:, 4] == 0)).mean()
    assert abs(both0 - np.exp(-(0.7 + 1.9))) < 3e-3


def test_small_state_spaces_share_the_count_call_between_four_rows():
    """S <= 8: the count uniform of row g is word g & 3 of the STREAM_JUMP_COUNT call of row group g >> 2 (the same
    per-row map as the Euler draw), the picks stay on the row's own call; rows that share a call draw independently."""
    n, off, seed = 200_000, 5, 77
    lam = np.tile(np.array([0.3, 0.0, 0.5], dtype=np.float32), (n, 1))
    k, K = rng.poisson_rows(lam, 8, off, seed)
    v0 = rng.row_units(n, 8, off, rng.STREAM_JUMP_COUNT, seed)
    assert np.array_equal(K, rng.poisson_from_unit(lam.sum(1, dtype=np.float32), v0))
    assert not np.array_equal(v0, rng.rowjump_total_unit(n, 8, off, seed))
    one = np.flatnonzero(K == 1)                                   # rows with one jump: pick 0 = word 1 of the row's own call 0
    pick = rng.rowjump_pick_units(np.arange(n, dtype=np.uint64)[one] + np.uint64(8), off, seed, 1)[:, 0]
    want = np.where(np.minimum(pick, np.float32(0.99999994)) * np.float32(0.8) < np.float32(0.3), 0, 2)
    assert np.array_equal(k[one].argmax(1), want)
    Kf = K.astype(np.float64)
    for d in (1, 2, 3):                                            # neighbours inside / across a call: uncorrelated counts
        assert abs(np.corrcoef(Kf[:-d], Kf[d:])[0, 1]) < 0.01
    quad = K[: n // 4 * 4].reshape(-1, 4)
    p0 = np.exp(-0.8)
    assert abs((quad == 0).all(1).mean() - p0 ** 4) < 4e-3         # P(all four rows of a call draw no jump)


def test_poisson_rows_is_a_function_of_global_row_and_offset():
    g = np.random.Generator(np.random.PCG64(5))
    lam = (g.random((64, 16)) * 0.4).astype(np.float32)
    a, _ = rng.poisson_rows(lam, 0, 3, 11)
    b0, _ = rng.poisson_rows(lam[:24], 0, 3, 11)
    b1, _ = rng.poisson_rows(lam[24:], 24, 3, 11)
    assert np.array_equal(a, np.concatenate([b0, b1]))          # batch sharding does not change the draws
    c, _ = rng.poisson_rows(lam, 0, 4, 11)
    assert not np.array_equal(a, c)                             # a new call offset gives new draws
    big = np.full((4, 8), 200.0, dtype=np.float32)              # total 1600 per row: picks are capped, counts bounded
    kb, Kb = rng.poisson_rows(big, 0, 0, 1)
    assert np.all(Kb > 1000) and np.all(kb.sum(1) <= rng.JUMP_PICK_CAP)


def test_poisson_from_unit_pmf_at_large_rates():
    """poisson_from_unit against scipy's Poisson law at lambda = 64 (the last rate on the exact pmf recurrence), 100 and
    400 (Cornish-Fisher quantile, a documented deviation from the reference's exact torch.poisson): on a uniform grid of
    400 001 quantiles the drawn counts reproduce mean and variance to 0.2 % and every probability P(K <= k) to 1.5e-3
    - which bounds the total-variation distance of the jump counts the samplers draw in the clamp-saturated regime."""
    from scipy.stats import poisson
    n = 400001
    v = ((np.arange(n) + 0.5) / n).astype(np.float32)
    for lam in (64.0, 100.0, 400.0):
        k = rng.poisson_from_unit(np.full(n, lam, dtype=np.float32), v)
        assert abs(k.mean() - lam) <= 2e-3 * lam, (lam, k.mean())
        assert abs(k.var() - lam) <= 1e-2 * lam, (lam, k.var())
        ks = np.arange(int(lam - 6 * lam ** 0.5), int(lam + 6 * lam ** 0.5))
        emp = np.searchsorted(np.sort(k), ks, side="right") / n          # P(K <= k) of the map
        assert np.abs(emp - poisson.cdf(ks, lam)).max() <= 1.5e-3, (lam, np.abs(emp - poisson.cdf(ks, lam)).max())
