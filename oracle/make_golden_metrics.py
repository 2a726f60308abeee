"""TEST INFRASTRUCTURE ONLY — pins oracle/metrics_oracle.py against the reference's own lib/datasets/metrics.py
(binary_hamming_sim, binary_exp_hamming_sim, binary_mmd, binary_exp_hamming_mmd) and writes tests/golden/metrics.npz.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_metrics
ml_collections (imported by lib.datasets.synthetic, unused by these functions) is absent here and replaced by an inert mock.
"""
from __future__ import annotations

import os
import sys
from unittest import mock

import numpy as np
import torch

from . import metrics_oracle as mo, ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, N, M, D, S, bandwidth, seed — C1 (binary, D=32) and C2 (maze, S=3, D=225) shapes, ragged tiles, a tiny set
CASES = [("mmd_c1", 300, 260, 32, 2, 0.1, 21), ("mmd_c2", 130, 97, 225, 3, 0.1, 22), ("mmd_bw", 65, 64, 32, 2, 0.7, 23),
         ("mmd_tiny", 2, 3, 5, 2, 0.1, 24), ("mmd_same", 128, 128, 40, 2, 0.1, 25)]


def import_reference_metrics():
    rh.import_reference()
    for _ in range(16):
        try:
            import lib.datasets.metrics as rm
            return rm
        except ImportError as e:
            if not e.name:
                raise
            sys.modules[e.name] = mock.MagicMock()
    raise RuntimeError("could not import lib.datasets.metrics")


def main():
    rm = import_reference_metrics()
    out = {}
    for name, N, M, D, S, bw, seed in CASES:
        x, y = mo.metric_inputs(seed, N, M, D, S)
        if name == "mmd_same":
            y = x.copy()
        tx, ty = torch.from_numpy(x), torch.from_numpy(y)
        k_ref = rm.binary_exp_hamming_sim(tx.float(), ty.float(), bw).numpy()
        h_ref = rm.binary_hamming_sim(tx.float(), ty.float()).numpy()
        mmd_ref = float(rm.binary_exp_hamming_mmd(tx, ty, None, bandwidth=bw))
        hmmd_ref = float(rm.binary_mmd(tx, ty, None, rm.binary_hamming_sim))
        k = mo.binary_exp_hamming_sim(x, y, bw)
        assert np.array_equal(mo.binary_hamming_sim(x, y), h_ref), name               # exact integers
        assert np.abs(k - k_ref).max() <= 2e-7 * max(1.0, k_ref.max()), name          # fp32 exp, libm vs torch: ~1 ulp
        got, hgot = mo.mmd(x, y, bw), mo.mmd(x, y, hamming=True)
        # the reference sums N*M fp32 values in fp32: its own distance from the fp64 value is the yardstick
        assert abs(got - mmd_ref) <= 5e-6 * max(k_ref.mean(), abs(mmd_ref)) + 1e-9, (name, got, mmd_ref)
        assert abs(hgot - hmmd_ref) <= 2e-5 * D, (name, hgot, hmmd_ref)
        out[f"{name}/sim"] = k_ref
        out[f"{name}/mmd_ref"] = np.float64(mmd_ref)
        out[f"{name}/mmd64"] = np.float64(got)
        out[f"{name}/hamming_mmd_ref"] = np.float64(hmmd_ref)
        out[f"{name}/hamming_mmd64"] = np.float64(hgot)
        out[f"{name}/sums64"] = mo.mmd_sums(x, y, bw)
        print(f"{name}: mmd ref {mmd_ref:.9g} oracle64 {got:.9g}; hamming ref {hmmd_ref:.7g} oracle64 {hgot:.7g}")
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)


if __name__ == "__main__":
    main()
