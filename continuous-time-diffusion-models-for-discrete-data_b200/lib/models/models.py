"""Output head of the image models: truncated logistic over S bins (reference lib/models/models.py:20-74, :248-282).

Only the head lives here — the score networks (U-Net, DiT, ...) stay the reference's stock PyTorch modules.  The
reference turns the two numbers per dimension the U-Net emits (`model_output == 'logistic_pars'`,
lib/networks/unet.py:450-452) into an (B, D, S) logits tensor with ~20 elementwise passes; here

* under `torch.no_grad()` (every sampler) `TruncatedLogisticHead.head_forward` returns an `ops.LogisticHead`, which the
  samplers pass straight into the fused reverse-step kernel: the logits never exist in memory;
* `sample_logistic(...)` is the reference's function with the same signature, backed by one CUDA pass
  (`ctdd_logistic_logits`) when no gradient is required;
* with gradients enabled (training) the head is one `autograd.Function`: the same forward kernel, and a backward kernel
  (`ctdd_logistic_logits_backward`) that recomputes the head from (mu, log_scale) and reads the incoming (B, D, S) gradient
  once — instead of autograd through ~20 saved (B, D, S) tensors.  `_logistic_logits_torch` keeps the same closed form in
  differentiable torch ops for host tensors / double precision checks.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ... import ops


def log_minus_exp(a, b, eps=1e-6):
    """log(exp(a) - exp(b)) for b < a with the reference's guard (lib/models/models.py:20-25)."""
    return a + torch.log1p(-torch.exp(b - a) + eps)


def _logistic_logits_autograd(mu, log_scale, S, fix_logistic):
    """Differentiable head on the training path: CUDA forward + backward kernels for fp32 CUDA tensors."""
    if mu.is_cuda and mu.dtype == torch.float32 and log_scale.dtype == torch.float32:
        return ops.logistic_logits_autograd(mu, log_scale, S, fix_logistic).view(*mu.shape, S)
    return _logistic_logits_torch(mu, log_scale, S, fix_logistic)


def _logistic_logits_torch(mu, log_scale, S, fix_logistic):
    """The head in differentiable torch ops: log(u_{s+1} (kappa v_s + 1e-6)) with u = sigmoid(z), v = sigmoid(-z) at the
    bin edges (identical in exact arithmetic to the reference's logsigmoid / log_minus_exp chain; see csrc/ctdd_head.cu)."""
    mu = mu.unsqueeze(-1)
    inv = torch.exp(2.0 - log_scale).unsqueeze(-1)
    edges = torch.linspace(-1.0, 1.0, S + 1, device=mu.device, dtype=mu.dtype)
    z = (edges - mu) * inv
    log_u, log_v = F.logsigmoid(z), F.logsigmoid(-z)
    kappa = -torch.expm1(-inv * (2.0 / S))
    logits = log_u[..., 1:] + torch.log(kappa * torch.exp(log_v[..., :-1]) + 1e-6)
    if fix_logistic:
        logits = torch.minimum(logits, log_v[..., :-1] + torch.log(kappa * torch.exp(log_u[..., 1:]) + 1e-6))
    return logits


def sample_logistic(net_out, B, C, D, S, fix_logistic, device):
    """Same signature and result layout as the reference's sample_logistic (lib/models/models.py:28-74):
    net_out = (mu, log_scale), each (B, C, H, W) -> logits (B, C, H, W, S)."""
    mu, log_scale = net_out[0], net_out[1]
    if torch.is_grad_enabled() and (mu.requires_grad or log_scale.requires_grad):
        return _logistic_logits_autograd(mu, log_scale, S, fix_logistic)
    return ops.logistic_logits(mu, log_scale, S, fix_logistic).view(*mu.shape, S)


class TruncatedLogisticHead:
    """Mixin for image x0-prediction models whose network returns (mu, log_scale) — the tail of
    ImageX0PredBasePaul.forward (reference lib/models/models.py:248-290).  Needs `self.S` and `self.fix_logistic`."""

    def head_forward(self, net_out, B, D):
        mu, log_scale = net_out[0], net_out[1]
        if torch.is_grad_enabled() and (mu.requires_grad or log_scale.requires_grad):
            return _logistic_logits_autograd(mu, log_scale, self.S, self.fix_logistic).view(B, D, self.S)
        return ops.LogisticHead(mu, log_scale, self.fix_logistic)
