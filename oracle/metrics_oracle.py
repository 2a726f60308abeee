"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's exp-Hamming MMD (lib/datasets/metrics.py:6-56) and of
the per-dimension histogram check.  Never imported by the product; pinned by oracle/make_golden_metrics.py against the
reference's own functions (tests/golden/metrics.npz).  Sums in fp64 (the reference sums N*M fp32 kernel values in fp32; the
fixture records how far its result sits from the fp64 value)."""
from __future__ import annotations

import numpy as np


def l1_distances(x, y):
    """(N, M) sum_d |x_d - y_d| — exact small integers.  metrics.py:7-9 / :15-18."""
    x = np.asarray(x, dtype=np.float32)
    y = np.asarray(y, dtype=np.float32)
    out = np.zeros((x.shape[0], y.shape[0]), dtype=np.float32)
    for d in range(x.shape[1]):
        out += np.abs(x[:, None, d] - y[None, :, d])
    return out


def binary_hamming_sim(x, y):
    return np.float32(np.asarray(x).shape[-1]) - l1_distances(x, y)


def binary_exp_hamming_sim(x, y, bd):
    """fp32 exp of the fp32 product, as torch.exp(-bd * d) computes it (metrics.py:22)."""
    return np.exp((np.float32(-bd) * l1_distances(x, y)).astype(np.float32)).astype(np.float32)


def mmd_sums(x, y, bd=0.1, hamming=False):
    """fp64 (sum_{i!=j} kxx, sum_{i!=j} kyy, sum kxy) of the fp32 kernel values."""
    sim = (lambda a, b: binary_hamming_sim(a, b)) if hamming else (lambda a, b: binary_exp_hamming_sim(a, b, bd))
    kxx = sim(x, x).astype(np.float64)
    kyy = sim(y, y).astype(np.float64)
    kxy = sim(x, y).astype(np.float64)
    return np.array([kxx.sum() - np.trace(kxx), kyy.sum() - np.trace(kyy), kxy.sum()])


def mmd(x, y, bd=0.1, hamming=False):
    """metrics.py:37-47 with the sums in fp64."""
    n, m = len(x), len(y)
    s = mmd_sums(x, y, bd, hamming)
    return s[0] / n / (n - 1) + s[1] / m / (m - 1) - 2 * s[2] / n / m


def state_histogram(x, S):
    x = np.asarray(x)
    out = np.zeros((x.shape[1], S), dtype=np.int32)
    for d in range(x.shape[1]):
        out[d] = np.bincount(x[:, d], minlength=S)
    return out


def metric_inputs(seed, N, M, D, S, shift=0.15):
    """Two sample sets with slightly different per-dimension state probabilities."""
    g = np.random.Generator(np.random.PCG64(seed))
    p = g.dirichlet(np.ones(S), size=D)
    q = (1 - shift) * p + shift * g.dirichlet(np.ones(S), size=D)
    x = np.stack([g.choice(S, size=N, p=p[d]) for d in range(D)], axis=1)
    y = np.stack([g.choice(S, size=M, p=q[d]) for d in range(D)], axis=1)
    return x.astype(np.int64), y.astype(np.int64)
