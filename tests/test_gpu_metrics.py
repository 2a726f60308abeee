"""GPU: exp-Hamming MMD kernels and the per-dimension histogram against the fixture the reference's own
lib/datasets/metrics.py produced and against the fp64 oracle."""
import functools

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as mo
from oracle.make_golden_metrics import CASES
from test_oracle_metrics import _inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_mmd_kernels_match_reference_fixture(golden, case):
    from ctdd_b200.lib.datasets import metrics
    from ctdd_b200 import ops
    g = golden["metrics"]
    name, N, M, D, S, bw, seed = case
    x, y = _inputs(case)
    tx, ty = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    K = metrics.binary_exp_hamming_sim(tx, ty, bw).cpu().numpy()
    ref = g[f"{name}/sim"]
    assert np.abs(K - ref).max() <= 2e-7 * max(1.0, ref.max())                    # fp32 expf vs torch.exp: ~1 ulp
    H = metrics.binary_hamming_sim(tx, ty).cpu().numpy()
    assert np.array_equal(H, mo.binary_hamming_sim(x, y))                          # exact integers
    sums = ops.pair_similarity_sums(tx, ty, bd=bw).cpu().numpy()
    assert np.allclose(sums, g[f"{name}/sums64"], rtol=2e-7, atol=1e-12)          # fp64 sums of ~1-ulp-different fp32 values
    mmd = metrics.binary_exp_hamming_mmd(tx, ty, None, bandwidth=bw)
    assert mmd.dtype == torch.float32 and mmd.dim() == 0 and mmd.is_cuda
    scale = max(ref.mean(), abs(float(g[f"{name}/mmd64"])))
    assert abs(float(mmd) - float(g[f"{name}/mmd64"])) <= 1e-6 * scale + 1e-9
    assert abs(float(mmd) - float(g[f"{name}/mmd_ref"])) <= 5e-6 * scale + 1e-9   # the reference's own fp32 summation noise
    hm = metrics.binary_hamming_mmd(tx, ty)
    assert abs(float(hm) - float(g[f"{name}/hamming_mmd64"])) <= 1e-6 * D
    # deterministic, and the generic-callable path (dense formula of the reference) agrees
    assert float(metrics.binary_exp_hamming_mmd(tx, ty, None, bandwidth=bw)) == float(mmd)
    dense = metrics.binary_mmd(tx, ty, None, lambda a, b: torch.exp(-bw * (a.unsqueeze(1) - b.unsqueeze(0)).abs().sum(-1)))
    assert abs(float(dense) - float(mmd)) <= 5e-6 * scale + 1e-9


def test_mmd_at_eval_synthetic_size_properties():
    """N = 4096, D = 32 (eval_synthetic.py:159): the reference would need a 2 GB (N, N, D) tensor.  Properties: symmetric
    in (x, y); row-permutation invariant; a shifted distribution gives a far larger value than a resample of the same one;
    sums of a set with itself relate as xy = xx + N; a 512-row slice equals the oracle."""
    from ctdd_b200.lib.datasets import metrics
    from ctdd_b200 import ops
    g = np.random.Generator(np.random.PCG64(9))
    N, D = 4096, 32
    x = torch.from_numpy((g.random((N, D)) < 0.5).astype(np.int64)).cuda()
    x2 = torch.from_numpy((g.random((N, D)) < 0.5).astype(np.int64)).cuda()
    y = torch.from_numpy((g.random((N, D)) < 0.6).astype(np.int64)).cuda()
    a, b = float(metrics.binary_exp_hamming_mmd(x, y)), float(metrics.binary_exp_hamming_mmd(y, x))
    assert a == pytest.approx(b, rel=1e-6)
    perm = torch.from_numpy(g.permutation(N)).cuda()
    assert float(metrics.binary_exp_hamming_mmd(x[perm], y)) == pytest.approx(a, rel=1e-5)
    same = float(metrics.binary_exp_hamming_mmd(x, x2))
    assert abs(same) < 0.1 * a
    s = ops.pair_similarity_sums(x, x).cpu().numpy()
    assert s[0] == s[1] and s[2] == pytest.approx(s[0] + N, rel=1e-12)            # the xy sum includes the diagonal (k = 1)
    # a 512-row slice against the oracle
    xs, ys = x[:512].cpu().numpy(), y[:512].cpu().numpy()
    assert float(metrics.binary_exp_hamming_mmd(x[:512], y[:512])) == pytest.approx(mo.mmd(xs, ys, 0.1), rel=1e-5, abs=1e-9)


def test_pair_similarity_empty_and_ragged():
    from ctdd_b200 import ops
    x = torch.zeros((0, 8), device="cuda")
    y = torch.ones((5, 8), device="cuda")
    assert ops.pair_similarity(x, y).shape == (0, 5)
    s = ops.pair_similarity_sums(x, y).cpu().numpy()
    assert s[0] == 0 and s[2] == 0 and s[1] == pytest.approx(20.0)                # 5*4 identical pairs, k = 1
    with pytest.raises(ValueError):
        ops.pair_similarity(y, torch.ones((5, 7), device="cuda"))


@pytest.mark.parametrize("shape", [(1000, 32, 2), (257, 225, 3), (3000, 48, 256), (5, 3, 8192), (1, 1, 2)])
def test_state_histogram_matches_bincount(shape):
    from ctdd_b200.lib.datasets import metrics
    N, D, S = shape
    g = np.random.Generator(np.random.PCG64(N + D))
    x = g.integers(0, S, (N, D))
    h = metrics.state_histograms(torch.from_numpy(x).cuda(), S)
    assert h.shape == (D, S) and np.array_equal(h.cpu().numpy(), mo.state_histogram(x, S))   # bit-exact
    h2 = metrics.state_histograms(torch.from_numpy(x).cuda(), S, counts=h.clone())
    assert np.array_equal(h2.cpu().numpy(), 2 * mo.state_histogram(x, S))
    bad = torch.from_numpy(x).cuda().clone()
    bad[0, 0] = S
    with pytest.raises(ValueError):
        metrics.state_histograms(bad, S)
