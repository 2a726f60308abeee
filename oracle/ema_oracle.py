"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's EMA update (lib/models/models.py:745-758).

Never imported by the product (ctdd_b200); used by tests/ as the checker of `ctdd_ema_update`.
Pinned by oracle/make_golden_ema.py against the reference's own `EMA.update_ema` run here (tests/golden/ema.npz).

    num_updates += 1
    decay = min(decay_cfg, (1 + num_updates) / (10 + num_updates))          # python floats (fp64)
    shadow <- shadow - fp32(1 - decay) * (shadow - param)                     # three fp32 roundings, no fma
"""
from __future__ import annotations

import numpy as np


def effective_decay(decay_cfg: float, num_updates: int) -> float:
    """Decay used by update number `num_updates` (1-based), models.py:750-752."""
    return min(decay_cfg, (1 + num_updates) / (10 + num_updates))


def ema_update(shadows, params, decay_cfg: float, num_updates: int):
    """One update of every shadow tensor; `num_updates` is the counter BEFORE the update. Returns (new shadows, counter)."""
    num_updates += 1
    omd = np.float32(1.0 - effective_decay(decay_cfg, num_updates))
    out = []
    for s, p in zip(shadows, params):
        s = np.asarray(s, dtype=np.float32)
        p = np.asarray(p, dtype=np.float32)
        diff = (s - p).astype(np.float32)
        out.append((s - (omd * diff).astype(np.float32)).astype(np.float32))
    return out, num_updates


def ema_inputs(seed: int, shapes, steps: int):
    """Deterministic parameter trajectories: params[k][i] is tensor i after optimiser step k (k = 0 is the initial value)."""
    g = np.random.Generator(np.random.PCG64(seed))
    traj = [[g.standard_normal(sh).astype(np.float32) for sh in shapes]]
    for _ in range(steps):
        traj.append([(p + np.float32(0.05) * g.standard_normal(p.shape).astype(np.float32)).astype(np.float32) for p in traj[-1]])
    return traj
