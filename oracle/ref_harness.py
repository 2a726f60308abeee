"""TEST INFRASTRUCTURE ONLY — drives the UNMODIFIED reference (read-only at /root/reference) with injected randomness.

Used only by oracle/make_golden.py (in the build container; /root/reference does not exist on the GPU box) to
pin oracle/ctmc_oracle.py and to generate tests/golden/*.npz.  Nothing here is imported by the product.

The reference needs `torchtyping` (absent) -> 5-line shim.  lib.models.models is NOT imported (pulls timm /
matplotlib); stub models compose nn.Module with the reference's own rate mixins instead.
Randomness is injected by patching, inside the imported reference only:
  torch.distributions.poisson.Poisson.sample       -> rng.poisson_rows(rate) (superposition map on per-row uniforms)
  torch.distributions.categorical.Categorical.sample -> rng.inv_cdf(probs, per-row uniforms of a scheduled stream)
  lib.sampling.sampling.get_initial_samples        -> oracle initial samples
  torch.rand (loss time draw)                      -> injected ts
"""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import rng
from . import ctmc_oracle as oc

import os

# the read-only reference in the build container; on the GPU box (no /root/reference) the copy that
# tools/stage_reference.py placed under the git-ignored baseline/_ref/ (bench.py --impl reference only)
REF_ROOT = "/root/reference/TAUnSDDM"
if not os.path.isdir(REF_ROOT):
    REF_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "TAUnSDDM")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "sampling"))


class Cfg(dict):
    """attribute-dict stand-in for ml_collections.ConfigDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def make_cfg(**sections) -> Cfg:
    c = Cfg()
    for k, v in sections.items():
        c[k] = Cfg(v) if isinstance(v, dict) else v
    return c


def import_reference():
    if "torchtyping" not in sys.modules:
        m = types.ModuleType("torchtyping")

        class TensorType:
            def __class_getitem__(cls, item):
                return cls

        m.TensorType = TensorType
        sys.modules["torchtyping"] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import lib.models.forward_model as fm
    import lib.models.model_utils as mu
    import lib.sampling.sampling as ss
    import lib.sampling.sampling_utils as su
    import lib.losses.losses as ll
    import lib.losses.losses_utils as lu
    return types.SimpleNamespace(fm=fm, mu=mu, ss=ss, su=su, ll=ll, lu=lu)


class StubNet(nn.Module):
    """logits[n,d,:] = A[x[n,d]] + t[n]*Bv[d] + w  — bitwise identical on CPU and CUDA (no fused ops)."""

    def __init__(self, S: int, D: int, seed: int, scale: float = 0.5, width: float = None):
        super().__init__()
        g = np.random.Generator(np.random.PCG64(seed))
        A = scale * g.standard_normal((S, S))
        if width is not None:  # denoiser-like: mass concentrated around the current state
            s = np.arange(S)
            A = A - (s[None, :] - s[:, None]) ** 2 / (2.0 * width ** 2)
        self.register_buffer("A", torch.from_numpy(A.astype(np.float32)))
        self.register_buffer("Bv", torch.from_numpy((scale * g.standard_normal((D, S))).astype(np.float32)))
        self.w = nn.Parameter(torch.zeros(S))

    def net(self, x, t):
        tb = t.view(-1, 1, 1).to(torch.float32) * self.Bv.unsqueeze(0)
        out = self.A[x.long()] + tb
        return out + self.w


class HeadStubNet(nn.Module):
    """(mu, log_scale) per dimension, shaped like the U-Net's `logistic_pars` output (lib/networks/unet.py:450-452):
    mu = tanh(loc + normalised input), log_scale growing with the noise level."""

    def __init__(self, S: int, D: int, seed: int):
        super().__init__()
        g = np.random.Generator(np.random.PCG64(seed))
        self.S_states = S
        self.register_buffer("loc_d", torch.from_numpy((0.2 * g.standard_normal(D)).astype(np.float32)))
        self.register_buffer("ls_d", torch.from_numpy((0.3 * g.standard_normal(D)).astype(np.float32)))
        self.w = nn.Parameter(torch.zeros(1))

    def head_params(self, x, t):
        tt = t.view(-1, 1).to(torch.float32)
        xn = (x.to(torch.float32) + 0.5) * (2.0 / self.S_states) - 1.0
        mu = torch.tanh(xn + self.loc_d * tt + self.w)
        log_scale = -1.2 + 2.0 * tt + self.ls_d + 0.0 * mu
        return mu, log_scale


def make_ref_model(ref, mixin_name: str, cfg, S: int, D: int, seed: int, scale: float = 0.5, width: float = None):
    mixin = getattr(ref.fm, mixin_name)

    class M(StubNet, mixin):
        def __init__(self):
            StubNet.__init__(self, S, D, seed, scale, width)
            mixin.__init__(self, cfg, "cpu")

        def forward(self, x, t):
            return self.net(x, t)

    return M()


class Injector:
    """Patches the reference's random draws; `schedule` lists the stream of each Categorical.sample call."""

    def __init__(self, ref, seed: int, init_x: torch.Tensor = None, ts: torch.Tensor = None,
                 cat_schedule=None, cat_offset: int = 0, cat_log: list = None):
        self.ref, self.seed, self.init_x, self.ts = ref, seed, init_x, ts
        self.call = 0                 # jump / Euler call counter (the kernels' `offset`)
        self.cat_schedule = list(cat_schedule) if cat_schedule else None
        self.cat_offset = cat_offset
        self.cat_i = 0
        self.cat_log = cat_log        # optional list collecting every Categorical draw (loss fixtures)

    def __enter__(self):
        P = torch.distributions.poisson.Poisson
        C = torch.distributions.categorical.Categorical
        self._saved = (P.sample, C.sample, self.ref.ss.get_initial_samples, torch.rand)
        inj = self

        def p_sample(self_, sample_shape=torch.Size()):
            rate = self_.rate
            S = rate.shape[-1]
            lam = rate.detach().numpy().astype(np.float32).reshape(-1, S)
            counts, _ = rng.poisson_rows(lam, 0, inj.call, inj.seed)
            inj.call += 1
            return torch.from_numpy(counts.reshape(rate.shape)).to(rate.dtype)

        def c_sample(self_, sample_shape=torch.Size()):
            probs = self_.probs.detach().numpy().astype(np.float32)
            shp = probs.shape[:-1]
            p2 = probs.reshape(-1, probs.shape[-1])
            if inj.cat_schedule is not None:
                stream = inj.cat_schedule[inj.cat_i % len(inj.cat_schedule)]
                inj.cat_i += 1
                off = inj.cat_offset
            else:
                stream, off = rng.STREAM_ROW, inj.call
                inj.call += 1
            v = rng.row_units(p2.shape[0], 0, off, stream, inj.seed)
            draw = torch.from_numpy(rng.inv_cdf(p2, v).reshape(shp))
            if inj.cat_log is not None:
                inj.cat_log.append(draw.clone())
            return draw

        def init_samples(N, D, device, S, initial_dist, initial_dist_std=None):
            if inj.init_x is not None:
                return inj.init_x.clone()
            return oc.initial_samples(N, D, S, initial_dist, initial_dist_std, inj.seed)

        real_rand = torch.rand

        def rand(*a, **k):
            if inj.ts is not None:
                shape = a[0] if (len(a) == 1 and isinstance(a[0], (tuple, list, torch.Size))) else a
                if tuple(shape) == tuple(inj.ts.shape):
                    return inj.ts.clone()
            return real_rand(*a, **k)

        P.sample, C.sample = p_sample, c_sample
        self.ref.ss.get_initial_samples = init_samples
        torch.rand = rand
        self._stdout = contextlib.redirect_stdout(open("/dev/null", "w"))
        self._stdout.__enter__()
        return self

    def __exit__(self, *exc):
        self._stdout.__exit__(*exc)
        P = torch.distributions.poisson.Poisson
        C = torch.distributions.categorical.Categorical
        P.sample, C.sample, self.ref.ss.get_initial_samples, torch.rand = self._saved
        return False
