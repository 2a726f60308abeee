"""CPU: oracle/metrics_oracle.py against the fixture the reference's own lib/datasets/metrics.py produced."""
import numpy as np
import pytest

from oracle import metrics_oracle as mo
from oracle.make_golden_metrics import CASES


def _inputs(case):
    name, N, M, D, S, bw, seed = case
    x, y = mo.metric_inputs(seed, N, M, D, S)
    if name == "mmd_same":
        y = x.copy()
    return x, y


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_metrics_oracle_matches_reference_fixture(golden, case):
    g = golden["metrics"]
    name, N, M, D, S, bw, seed = case
    x, y = _inputs(case)
    k = mo.binary_exp_hamming_sim(x, y, bw)
    ref = g[f"{name}/sim"]
    assert k.shape == ref.shape == (N, M)
    assert np.abs(k - ref).max() <= 2e-7 * max(1.0, ref.max())
    got = mo.mmd(x, y, bw)
    assert got == pytest.approx(float(g[f"{name}/mmd64"]), rel=1e-9, abs=1e-15)
    # distance to the reference's fp32-summed value: within the reference's own summation noise
    assert abs(got - float(g[f"{name}/mmd_ref"])) <= 5e-6 * max(ref.mean(), abs(got)) + 1e-9
    assert abs(mo.mmd(x, y, hamming=True) - float(g[f"{name}/hamming_mmd_ref"])) <= 2e-5 * D


def test_histogram_oracle_and_kl_host_logic():
    import torch
    from ctdd_b200.lib.datasets import metrics
    x, y = mo.metric_inputs(3, 500, 400, 7, 5)
    hx, hy = mo.state_histogram(x, 5), mo.state_histogram(y, 5)
    assert hx.sum() == 500 * 7 and (hx.sum(1) == 500).all()
    kl = metrics.histogram_kl(torch.from_numpy(hx), torch.from_numpy(hy))
    assert kl.shape == (7,) and float(kl.min()) >= 0
    assert float(metrics.histogram_kl(torch.from_numpy(hx), torch.from_numpy(hx)).max()) == 0.0


def test_metrics_have_no_cpu_fallback():
    import torch
    from ctdd_b200.lib.datasets import metrics
    x = torch.zeros(4, 8)
    with pytest.raises(RuntimeError):
        metrics.binary_exp_hamming_mmd(x, x, None)
    with pytest.raises(RuntimeError):
        metrics.state_histograms(x.long(), 2)
