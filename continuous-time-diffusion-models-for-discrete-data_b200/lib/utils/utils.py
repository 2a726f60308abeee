"""Small tensor helpers with the semantics of the reference's lib/utils/utils.py (:59-62 expand_dims, :86-91 log1mexp)."""
import torch


def expand_dims(x, axis):
    """successive unsqueeze over `axis` (reference utils.py:59-62)."""
    for i in axis:
        x = x.unsqueeze(i)
    return x


def log1mexp(x):
    """log(1 - exp(-|x|)) with the -0.693 switch between log(-expm1) and log1p(-exp) (reference utils.py:86-91)."""
    x = -torch.abs(x)
    return torch.where(x > -0.693, torch.log(-torch.expm1(x)), torch.log1p(-torch.exp(x)))


def binary_exp_hamming_sim(x, y, bd):
    """(N, M) exp(-bd * L1 distance), reference utils.py:101-105 — one CUDA kernel, no (N, M, D) tensor."""
    from ..datasets import metrics
    return metrics.binary_exp_hamming_sim(x, y, bd)


def binary_exp_hamming_mmd(x, y, bandwidth=0.1):
    """Reference utils.py:127-129 (the cfg-less twin of lib/datasets/metrics.py:51-53)."""
    from ..datasets import metrics
    return metrics.binary_exp_hamming_mmd(x, y, None, bandwidth)
