"""TEST INFRASTRUCTURE ONLY — pins oracle/ema_oracle.py against the reference's own EMA mixin
(lib/models/models.py:729-826) and writes tests/golden/ema.npz.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_ema
The reference's `EMA` is mixed into a small nn.Module exactly as its model classes do (`class M(EMA, Net)`), its
parameters are driven along a fixed trajectory and `update_ema()` is called after every step; the shadow parameters after
each update are stored.  The restatement must reproduce them bit for bit.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ema_oracle as eo
from .make_golden_head import import_reference_head

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# ragged on purpose: odd sizes (vector tail), a tensor larger than one chunk, a scalar-like tensor, a frozen one
SHAPES = [(7,), (33, 5), (1,), (40000,), (64, 3, 3, 3), (129,)]
CASES = [("ema_9999", 0.9999, 14, 11), ("ema_05", 0.5, 6, 12), ("ema_0", 0.0, 3, 13), ("ema_1", 1.0, 3, 14)]


def main():
    _, mm = import_reference_head()
    from ctdd_b200.config import make_config
    out = {}
    for name, decay, steps, seed in CASES:
        traj = eo.ema_inputs(seed, SHAPES, steps)

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.from_numpy(p.copy())) for p in traj[0]])
                self.frozen = torch.nn.Parameter(torch.ones(5), requires_grad=False)

        class M(mm.EMA, Net):
            def __init__(self, cfg):
                mm.EMA.__init__(self, cfg)
                Net.__init__(self)
                self.init_ema()

        m = M(make_config(model=dict(ema_decay=decay), device="cpu"))
        shadows = [p.copy() for p in traj[0]]
        n = 0
        for k in range(1, steps + 1):
            with torch.no_grad():
                for p, v in zip(m.ps, traj[k]):
                    p.copy_(torch.from_numpy(v))
            m.update_ema()
            shadows, n = eo.ema_update(shadows, traj[k], decay, n)
            for i, (a, b) in enumerate(zip(m.shadow_params, shadows)):
                assert np.array_equal(a.numpy(), b), (name, k, i)        # restatement == reference, bitwise
        assert m.num_updates == n == steps
        sd = m.state_dict()
        assert sd["ema_num_updates"] == steps and sd["ema_decay"] == decay
        for i, s in enumerate(m.shadow_params):
            out[f"{name}/shadow{i}"] = s.numpy().copy()
        out[f"{name}/meta"] = np.array([decay, steps, seed], dtype=np.float64)
        print(name, "ok:", steps, "updates, oracle bitwise equal to the reference")
    np.savez_compressed(os.path.join(OUT, "ema.npz"), **out)


if __name__ == "__main__":
    main()
