"""Timing of the auxiliary kernels on a B200 (CUDA events, after warm-up): multi-tensor EMA against the reference's
per-parameter torch loop, exp-Hamming MMD against the reference's dense (N, N, D) formula, per-dimension histogram.
Prints one JSON line per kernel.   python tools/aux_probe.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctdd_b200 import ops  # noqa: E402
from ctdd_b200.lib.datasets import metrics  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def ema():
    # parameter-size profile of a CIFAR10-class U-Net: ~36 M parameters in ~1100 tensors (convs, norms, biases)
    g = np.random.Generator(np.random.PCG64(0))
    sizes = []
    for ch in (128, 256, 256, 256):
        for _ in range(40):
            sizes += [ch * ch * 9 // (4 if ch > 128 else 1), ch, ch, ch]
    sizes += [int(s) for s in g.integers(64, 4096, 400)]
    params = [torch.randn(n, device="cuda") for n in sizes]
    shadows = [p.clone() for p in params]
    ref_shadows = [p.clone() for p in params]
    for p in params:
        p.add_(0.01 * torch.randn_like(p))
    table = ops.EmaTable(shadows, params)
    omd = 1.0 - 0.9999

    def ref_loop():
        for s, p in zip(ref_shadows, params):
            s.sub_(omd * (s - p))

    def foreach():
        d = torch._foreach_sub(ref_shadows, params)
        torch._foreach_mul_(d, omd)
        torch._foreach_sub_(ref_shadows, d)

    table.update(omd)
    ref_loop()
    same = all(torch.equal(a, b) for a, b in zip(shadows, ref_shadows))
    t_k, t_ref, t_fe = timed(lambda: table.update(omd)), timed(ref_loop, reps=5), timed(foreach, reps=5)
    n = sum(sizes)
    print(json.dumps({"kernel": "ema_update_kernel", "tensors": len(sizes), "params": n, "chunks": table.n_chunks,
                      "ms": t_k, "GB/s": 12 * n / t_k / 1e6, "torch_loop_ms": t_ref, "torch_foreach_ms": t_fe,
                      "bitwise_equal_to_torch_loop": same}))


def mmd():
    g = np.random.Generator(np.random.PCG64(1))
    for N, D, S in ((4096, 32, 2), (1024, 225, 3), (16384, 32, 2)):
        x = torch.from_numpy(g.integers(0, S, (N, D))).cuda()
        y = torch.from_numpy(g.integers(0, S, (N, D))).cuda()
        t_k = timed(lambda: metrics.binary_exp_hamming_mmd(x, y), reps=10)
        rec = {"kernel": "pair_kernel<sum> x3 + finish", "N": N, "D": D, "ms": t_k,
               "pair_dims_per_s": 2 * N * N * D / t_k / 1e-3 / 1e12, "unit": "T pair-dims/s (upper triangles + cross)"}
        if N * N * D * 4 <= 3 << 30:
            xf, yf = x.float(), y.float()

            def dense():
                def sim(a, b):
                    return torch.exp(-0.1 * (a.unsqueeze(1) - b.unsqueeze(0)).abs().sum(-1))
                n = xf.shape[0]
                eye = 1 - torch.eye(n, device="cuda")
                return (sim(xf, xf) * eye).sum() / n / (n - 1) + (sim(yf, yf) * eye).sum() / n / (n - 1) - 2 * sim(xf, yf).sum() / n / n
            rec["torch_dense_ms"] = timed(dense, reps=3, warm=1)
            rec["abs_diff_vs_dense"] = abs(float(dense()) - float(metrics.binary_exp_hamming_mmd(x, y)))
        print(json.dumps(rec))


def hist():
    g = np.random.Generator(np.random.PCG64(2))
    for N, D, S in ((16384, 3072, 256), (16384, 32, 2)):
        x = torch.from_numpy(g.integers(0, S, (N, D)).astype(np.int32)).cuda()
        buf = torch.zeros(D * S + 1, dtype=torch.int32, device="cuda")
        from ctdd_b200 import _native as nat
        t = timed(lambda: nat.check(nat.lib().ctdd_state_histogram(x.data_ptr(), N, D, S, buf.data_ptr(), nat.stream())))
        print(json.dumps({"kernel": "histogram_kernel", "N": N, "D": D, "S": S, "ms": t, "GB/s": 4 * N * D / t / 1e6}))


if __name__ == "__main__":
    ema()
    mmd()
    hist()
