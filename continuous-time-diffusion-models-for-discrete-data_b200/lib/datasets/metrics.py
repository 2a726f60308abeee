"""Sample-quality metrics of the synthetic / maze evaluation — the exp-Hamming MMD family of the reference
(lib/datasets/metrics.py:6-56, :150-222; lib/utils/utils.py:101-129), same names and signatures.

The reference forms the (N, M, D) tensor of pairwise differences (2 GB at eval_synthetic's 4096 samples); here the pair
distances, the kernel value and the three sums of the MMD are one CUDA kernel each (`ctdd_pair_similarity[_sum]`), with
fp64 accumulation in a fixed order.  `state_histograms` / `histogram_kl` are the per-dimension histogram check used for
the distributional parity of the reverse process."""
import functools

import torch

from ... import ops


def binary_hamming_sim(x, y):
    """(N, M): D - sum_d |x_d - y_d|   (reference metrics.py:6-10)."""
    return ops.pair_similarity(x, y, hamming=True)


def binary_exp_hamming_sim(x, y, bd):
    """(N, M): exp(-bd * sum_d |x_d - y_d|)   (reference metrics.py:14-22)."""
    return ops.pair_similarity(x, y, bd=bd)


exp_hamming_sim = binary_exp_hamming_sim          # reference metrics.py:150-154 is the same function


def _mmd_from_sums(sums, n, m):
    kxx = sums[0] / n / (n - 1)
    kyy = sums[1] / m / (m - 1)
    kxy = sums[2] / n / m
    return (kxx + kyy - 2 * kxy).to(torch.float32)


def binary_mmd(x, y, cfg, sim_fn):
    """MMD of two sample sets under `sim_fn` (reference metrics.py:25-48).  The two similarity functions of this module
    are summed inside the kernel; any other callable gets the reference's dense formula."""
    f = sim_fn.func if isinstance(sim_fn, functools.partial) else sim_fn
    if f is binary_exp_hamming_sim or f is binary_hamming_sim:
        bd = 0.0
        if isinstance(sim_fn, functools.partial):
            bd = sim_fn.keywords.get("bd", sim_fn.args[0] if sim_fn.args else 0.0)
        sums = ops.pair_similarity_sums(x, y, bd=bd, hamming=f is binary_hamming_sim)
        return _mmd_from_sums(sums, x.shape[0], y.shape[0])
    x = x.to(torch.float32)
    y = y.to(torch.float32)
    n, m = x.shape[0], y.shape[0]
    kxx = sim_fn(x, x)
    kxx = torch.sum(kxx * (1 - torch.eye(n, device=x.device))) / n / (n - 1)
    kyy = sim_fn(y, y)
    kyy = torch.sum(kyy * (1 - torch.eye(m, device=x.device))) / m / (m - 1)
    kxy = torch.sum(sim_fn(x, y)) / n / m
    return kxx + kyy - 2 * kxy


def binary_exp_hamming_mmd(x, y, cfg=None, bandwidth=0.1):
    """Reference metrics.py:51-53 (cfg is unused there as well; lib/utils/utils.py:127-129 is the cfg-less twin)."""
    return binary_mmd(x, y, cfg, functools.partial(binary_exp_hamming_sim, bd=bandwidth))


def binary_hamming_mmd(x, y):
    """Reference metrics.py:55-56 (which forgets the cfg argument and would raise; this one works)."""
    return binary_mmd(x, y, None, binary_hamming_sim)


def eval_mmd(config, model, sampler, dataloader, n_rounds: int = 10, n_samples: int = 1024):
    """Average exp-Hamming MMD between `n_samples` data rows and `n_samples` generated rows over `n_rounds` rounds
    (reference metrics.py:168-222).  `sampler` is anything with `sample(model, N)` (the CTMC samplers of this package) or
    the reference's D3PM object with `p_sample_loop`."""
    n_data = max(1, n_samples // config.data.batch_size)
    total = None
    with torch.no_grad():
        for _ in range(n_rounds):
            gt = []
            while len(gt) < n_data:
                for batch in dataloader:
                    gt.append(batch)
                    if len(gt) == n_data:
                        break
            gt = torch.stack(gt, dim=0).view(-1, config.model.concat_dim).to(config.device)
            if hasattr(sampler, "p_sample_loop"):
                x0 = sampler.p_sample_loop(model, (n_samples, config.model.concat_dim), config.sampler.num_steps)
            else:
                x0 = sampler.sample(model, n_samples)
                x0 = x0[0] if isinstance(x0, tuple) else x0
            x0 = torch.as_tensor(x0).to(config.device)
            mmd = binary_exp_hamming_mmd(gt, x0, config)
            total = mmd if total is None else total + mmd
    return total / n_rounds


def state_histograms(x, S, counts=None):
    """Per-dimension state counts (D, S) of integer samples (N, D) — accumulates into `counts` when given."""
    return ops.state_histogram(torch.as_tensor(x), S, counts)


def histogram_kl(counts_p, counts_q, alpha=0.5):
    """Per-dimension symmetrised KL between two (D, S) count tables with add-alpha smoothing -> (D,) float64."""
    p = counts_p.to(torch.float64) + alpha
    q = counts_q.to(torch.float64) + alpha
    p = p / p.sum(-1, keepdim=True)
    q = q / q.sum(-1, keepdim=True)
    return 0.5 * ((p * (p / q).log()).sum(-1) + (q * (q / p).log()).sum(-1))
