"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's truncated-logistic output head.

Follows `sample_logistic` (lib/models/models.py:28-74) and the identical inline copy in `ImageX0PredBasePaul.forward`
(:248-282): the network emits (mu, log_scale) per dimension; the logit of state s is the log-mass a logistic
distribution N_logistic(mu, exp(log_scale - 2)) puts on bin s of [-1, 1] cut into S equal bins, computed as
log(exp(a) - exp(b)) with a 1e-6 guard inside the log1p (`log_minus_exp`, :20-25).  `fix_logistic` takes the minimum
with the mirrored evaluation.  Pinned against the reference's own function by oracle/make_golden_head.py ->
tests/golden/head.npz.  dtype-generic so the same code is the fp64 yardstick.  Nothing here is imported by the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _log_minus_exp(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """models.py:20-25: log(exp(a) - exp(b)) for b < a, guarded."""
    return a + torch.log1p(-torch.exp(b - a) + eps)


def truncated_logistic_logits(mu: torch.Tensor, log_scale: torch.Tensor, S: int, fix_logistic: bool) -> torch.Tensor:
    """mu, log_scale: (...,) -> logits (..., S) in the dtype of mu."""
    mu = mu.unsqueeze(-1)
    log_scale = log_scale.unsqueeze(-1)
    inv_scale = torch.exp(-(log_scale - 2))
    bin_width = 2.0 / S
    centers = torch.linspace(-1.0 + bin_width / 2, 1.0 - bin_width / 2, S, dtype=mu.dtype)
    left = (centers - bin_width / 2 - mu) * inv_scale
    right = (centers + bin_width / 2 - mu) * inv_scale
    left_logcdf = F.logsigmoid(left)
    right_logcdf = F.logsigmoid(right)
    logits = _log_minus_exp(right_logcdf, left_logcdf)
    if fix_logistic:
        logits = torch.minimum(logits, _log_minus_exp(-left + left_logcdf, -right + right_logcdf))
    return logits


def head_inputs(rows: int, seed: int, scale_lo: float = -3.0, scale_hi: float = 3.0):
    """Synthetic head parameters shaped like the U-Net's output (lib/networks/unet.py:450-452): mu = tanh(.) in (-1, 1),
    log_scale unconstrained (a trained network sits around -1 .. 2).  Returns fp32 (mu, log_scale) of shape (rows,)."""
    g = torch.Generator().manual_seed(seed)
    mu = torch.tanh(1.2 * torch.randn(rows, generator=g))
    log_scale = scale_lo + (scale_hi - scale_lo) * torch.rand(rows, generator=g)
    return mu.float(), log_scale.float()
