"""TEST INFRASTRUCTURE ONLY — CPU restatement of the forward noising / loss terms of lib/losses/losses.py."""
from __future__ import annotations

import numpy as np
import torch

from . import rng


def noise_xt(Q: torch.Tensor, R: torch.Tensor, x0: torch.Tensor, seed: int, offset: int, batch_offset: int = 0):
    """losses.py:46-101: x_t ~ Cat(Q[b, x0, :]); d* ~ Cat(sum_{s != x_t} R[b, x_t, s]); value ~ Cat(R[b, x_t[d*], .] off-diag).
    Q, R: (B,S,S); x0 (B,D). Uniforms: one per (b,d) for x_t (global row = (batch_offset+b)*D + d), one per b each for
    the dimension and the value."""
    B, D = x0.shape
    S = Q.shape[-1]
    rows = Q[torch.arange(B).repeat_interleave(D), x0.flatten().long(), :].numpy().astype(np.float32)
    v = rng.row_units(B * D, batch_offset * D, offset, rng.STREAM_NOISE_XT, seed)
    xt = rng.inv_cdf(rows, v).reshape(B, D)
    Rn = R.numpy().astype(np.float32)
    rv = Rn[np.arange(B)[:, None], xt, :].copy()                      # (B,D,S) rows R[b, x_t, :]
    rv[np.arange(B)[:, None], np.arange(D)[None, :], xt] = 0.0
    w = np.zeros((B, D), dtype=np.float32)
    for s in range(S):                                                # sequential fp32 sum over s (device order)
        w = (w + rv[:, :, s]).astype(np.float32)
    v1 = rng.row_units(B, batch_offset, offset, rng.STREAM_TILDE_DIM, seed)
    dstar = rng.inv_cdf(w, v1)
    v2 = rng.row_units(B, batch_offset, offset, rng.STREAM_TILDE_VAL, seed)
    newval = rng.inv_cdf(rv[np.arange(B), dstar, :], v2)
    xtil = xt.copy()
    xtil[np.arange(B), dstar] = newval
    return torch.from_numpy(xt), torch.from_numpy(xtil)
