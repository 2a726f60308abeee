"""Forward CTMC rate mixins — drop-in for the reference's lib/models/forward_model.py.

Same class names, constructor signature `(cfg, device)`, attributes and methods
(`rate`, `transition`, `rate_mat`, `transit_between`, `_rate_scalar`, `_integral_rate_scalar`) as
TAUnSDDM/lib/models/forward_model.py:9-306, so `class Model(EMA, Net, GaussianTargetRate)` compositions keep
working.  What changed: q_{t|0} is built by the batched-t CUDA kernel `ctdd_build_qt0`
(csrc/ctdd_qt0.cu) from the cached eigendecomposition instead of two batched matmuls + diag_embed, and the
samplers/losses ask for ONE matrix per distinct time (`qt0_tables`) instead of N identical copies.

The eigen tensors stay plain attributes (not buffers) so no new state_dict keys appear
(reference: forward_model.py:241-244; EMA.load_state_dict rejects unknown keys, lib/models/models.py:793-800).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from ... import _native as nat
from ..utils import utils


def _gaussian_target_base_rate(S: int, rate_sigma: float, Q_sigma: float) -> np.ndarray:
    """fp64 base rate of GaussianTargetRate (reference forward_model.py:216-236), vectorised.

    For i<j exactly one of R[i,j] (i+j < S) / R[j,i] (i+j >= S) is seeded with exp(-(|i-j|-1)^2/rate_sigma^2);
    the partner entry is the seeded one times the detailed-balance factor F of a Gaussian centred at S/2.
    """
    idx = np.arange(S)
    i, j = idx[:, None], idx[None, :]
    vals = np.exp(-np.arange(0, S, dtype=np.float64) ** 2 / (rate_sigma ** 2))
    seed = np.zeros((S, S))
    upper = (i < S // 2) & (j > i) & (j < S - i)
    lower = (i > S // 2) & (j < i) & (j > S - 1 - i)
    seed[upper] = vals[(j - i - 1)[upper]]
    seed[lower] = vals[(i - j - 1)[lower]]
    F = np.exp(-((j + 1.0) ** 2 - (i + 1.0) ** 2 + S * (i + 1.0) - S * (j + 1.0)) / (2 * Q_sigma ** 2))
    rate = seed.copy()
    fill = seed.T > 0.0                      # R[i,j] <- R[j,i] * F[i,j] wherever the partner is seeded
    rate[fill] = (seed.T * F)[fill]
    # the reference's in-place row-major sweep revisits a seeded lower entry after its upper partner was
    # filled and multiplies once more (by F[j,i] = 1/F[i,j]); replay that so the fp64 bits agree.
    again = (seed > 0.0) & (i > j)
    rate[again] = (rate.T * F)[again]
    rate = rate - np.diag(np.diag(rate))
    rate = rate - np.diag(np.sum(rate, axis=1))
    return rate


class _EigenForward:
    """Shared machinery: device copies of (R_b, U, U^-1, lambda) and the kernel-backed q_{t|0} builder."""

    _normalize = True       # divide rows by their sum before the clamp
    _clamp_below = 1e-8

    def _setup_eigen(self, rate: np.ndarray, eigvals: np.ndarray, eigvecs: np.ndarray, inv: np.ndarray, device):
        dev = torch.device(device)
        f = lambda a: torch.from_numpy(np.ascontiguousarray(np.real(a))).float().to(dev)
        self._Rb = f(rate)
        self._RbT = f(rate.T)
        self._lam = f(eigvals)
        self._U = f(eigvecs)
        self._Uinv = f(inv)
        self._qt0_cache = {}

    # --- kernel-backed builders -------------------------------------------------------------------------
    def _build_qt0(self, d_int: torch.Tensor, inverse: bool = True, want_transpose: bool = False):
        """Q_b = U diag(exp(lam * d_int[b])) U^-1, normalise, clamp. d_int: (B,) fp32 CUDA."""
        d_int = d_int.detach().to(torch.float32).contiguous()
        if not d_int.is_cuda:
            raise RuntimeError("ctdd_b200 forward models compute q_{t|0} on the GPU only; pass CUDA time tensors "
                               "(no CPU fallback)")
        if self._U.device != d_int.device:
            self._move_to(d_int.device)
        B, S = d_int.shape[0], self.S
        Q = torch.empty((B, S, S), dtype=torch.float32, device=d_int.device)
        QT = torch.empty_like(Q) if want_transpose else None
        right = self._Uinv if inverse else self._U_T()
        with nat.on_device(d_int.device):
            nat.check(nat.lib().ctdd_build_qt0(nat.ptr(self._U), nat.ptr(right), nat.ptr(self._lam), nat.ptr(d_int), B, S,
                                               1 if self._normalize else 0, self._clamp_below, nat.ptr(Q), nat.ptr(QT),
                                               nat.stream()), "ctdd_build_qt0")
        return (Q, QT) if want_transpose else Q

    def _U_T(self):
        if not hasattr(self, "_UT_cached") or self._UT_cached.device != self._U.device:
            self._UT_cached = self._U.t().contiguous()
        return self._UT_cached

    def _move_to(self, device):
        for name in ("_Rb", "_RbT", "_lam", "_U", "_Uinv"):
            setattr(self, name, getattr(self, name).to(device))
        self._qt0_cache = {}

    def _transition_delta(self, t: torch.Tensor) -> torch.Tensor:
        """integral of beta over [0, t] as the reference evaluates it (fp32 torch ops)."""
        return self._integral_rate_scalar(t)

    def qt0_tables(self, ts, device):
        """Q and Q^T for a list of DISTINCT times (host floats): ((T,S,S), (T,S,S)) fp32 CUDA, cached.

        The reference samplers call transition(t * ones(N)) and get N identical (S,S) matrices per step
        (sampling.py:34-35, quirk B.10); the kernels need one.  Times go through the same fp32 torch
        arithmetic as `t * torch.ones((N,))` so the matrices match the reference's.
        """
        key = (tuple(float(np.float32(t)) for t in ts), str(device))
        hit = self._qt0_cache.get(key)
        if hit is not None:
            return hit
        t32 = torch.tensor([float(t) for t in ts], dtype=torch.float64).to(torch.float32)
        d_int = self._transition_delta(t32).to(torch.float32)
        beta = self._rate_scalar(t32).to(torch.float32)
        Q, QT = self._build_qt0(d_int.to(device), inverse=True, want_transpose=True)
        out = (Q, QT, [float(b) for b in beta])
        if len(self._qt0_cache) > 8:
            self._qt0_cache.clear()
        self._qt0_cache[key] = out
        return out

    def base_rate_tables(self, device):
        """(R_b, R_b^T) fp32 CUDA, (S,S) each."""
        if self._Rb.device != torch.device(device):
            self._move_to(torch.device(device))
        return self._Rb, self._RbT

    def _scaled_rate(self, beta: torch.Tensor) -> torch.Tensor:
        beta = beta.detach().to(torch.float32).contiguous()
        if not beta.is_cuda:
            raise RuntimeError("ctdd_b200 forward models run on the GPU only (no CPU fallback)")
        if self._Rb.device != beta.device:
            self._move_to(beta.device)
        B, S = beta.shape[0], self.S
        out = torch.empty((B, S, S), dtype=torch.float32, device=beta.device)
        with nat.on_device(beta.device):
            nat.check(nat.lib().ctdd_build_rate(nat.ptr(self._Rb), nat.ptr(beta), B, S, nat.ptr(out), nat.stream()),
                      "ctdd_build_rate")
        return out


class BirthDeathForwardBase(_EigenForward):
    """reference forward_model.py:9-75."""

    def __init__(self, cfg, device):
        self.S = S = cfg.data.S
        self.sigma_min, self.sigma_max = cfg.model.sigma_min, cfg.model.sigma_max
        self.device = device
        base_rate = np.diag(np.ones((S - 1,)), 1) + np.diag(np.ones((S - 1,)), -1)
        base_rate -= np.diag(np.sum(base_rate, axis=1))
        eigvals, eigvecs = np.linalg.eigh(base_rate)
        self._setup_eigen(base_rate, eigvals, eigvecs, eigvecs.T, device)
        self.base_rate, self.base_eigvals, self.base_eigvecs = self._Rb, self._lam, self._U

    def _rate_scalar(self, t):
        return (self.sigma_min ** 2 * (self.sigma_max / self.sigma_min) ** (2 * t)
                * math.log(self.sigma_max / self.sigma_min))

    def _integral_rate_scalar(self, t):
        return 0.5 * self.sigma_min ** 2 * (self.sigma_max / self.sigma_min) ** (2 * t) - 0.5 * self.sigma_min ** 2

    def rate(self, t):
        return self._scaled_rate(self._rate_scalar(t))

    def transition(self, t):
        return self._build_qt0(self._integral_rate_scalar(t))


class UniformRate(_EigenForward):
    """reference forward_model.py:78-129: R = c(11^T - S I); transition is clamped but NOT renormalised."""

    _normalize = False

    def __init__(self, cfg, device):
        self.S = S = cfg.data.S
        self.rate_const = cfg.model.rate_const
        self.device = device
        rate = self.rate_const * np.ones((S, S))
        rate = rate - np.diag(np.diag(rate))
        rate = rate - np.diag(np.sum(rate, axis=1))
        eigvals, eigvecs = np.linalg.eigh(rate)
        self._setup_eigen(rate, eigvals, eigvecs, eigvecs.T, device)
        self.rate_matrix, self.eigvals, self.eigvecs = self._Rb, self._lam, self._U

    def _rate_scalar(self, t):
        return torch.ones_like(t)

    def _integral_rate_scalar(self, t):
        return t

    def rate(self, t):
        return self._scaled_rate(torch.ones_like(t, dtype=torch.float32))

    def rate_mat(self, y, t):
        del t
        return self._Rb.to(y.device)[y.long()]

    def transition(self, t):
        return self._build_qt0(t)

    def transit_between(self, t1, t2):
        return self.transition(t2 - t1)


class UniformVariantRate(UniformRate):
    """reference forward_model.py:132-204: same R, time-warped; transition renormalises, then clamps."""

    _normalize = True

    def __init__(self, config, device):
        super().__init__(config, device)
        self.config = config
        self.t_func = config.model.t_func
        self.device = config.device
        if self.t_func == "log":
            self.time_base = config.model.time_base
            self.time_exp = config.model.time_exp

    def _integral_rate_scalar(self, t):
        if self.t_func == "log_sqr":
            return torch.log(t ** 2 + 1)
        if self.t_func == "sqrt_cos":
            return -torch.sqrt(torch.cos(torch.pi / 2 * t))
        if self.t_func == "log":
            return self.time_base * (self.time_exp ** t) - self.time_base
        raise ValueError("Unknown t_func %s" % self.t_func)

    def _rate_scalar(self, t):
        if self.t_func == "log_sqr":
            return 2 * t / (t ** 2 + 1)
        if self.t_func == "sqrt_cos":
            t = torch.pi / 2 * t
            return torch.pi / 4.0 * (torch.sin(t) / torch.sqrt(torch.cos(t)))
        if self.t_func == "log":
            return self.time_base * math.log(self.time_exp) * self.time_exp ** t
        raise ValueError("Unknown t_func %s" % self.t_func)

    def _transition_delta(self, t):
        return self._integral_rate_scalar(t) - self._integral_rate_scalar(torch.zeros_like(t))

    def rate(self, t):
        return self._scaled_rate(self._rate_scalar(t))

    def rate_mat(self, y, t):
        r = self.rate(t)
        bidx = utils.expand_dims(torch.arange(t.size(0), device=r.device), axis=tuple(range(1, y.dim())))
        return r[bidx, y.long()]

    def transit_between(self, t1, t2):
        return self._build_qt0(self._integral_rate_scalar(t2) - self._integral_rate_scalar(t1))

    def transition(self, t):
        return self.transit_between(torch.zeros_like(t), t)


class GaussianTargetRate(_EigenForward):
    """reference forward_model.py:207-306."""

    def __init__(self, cfg, device):
        self.S = S = cfg.data.S
        self.rate_sigma = cfg.model.rate_sigma
        self.Q_sigma = cfg.model.Q_sigma
        self.time_exp = cfg.model.time_exp
        self.time_base = cfg.model.time_base
        self.device = device
        rate = _gaussian_target_base_rate(S, self.rate_sigma, self.Q_sigma)
        eigvals, eigvecs = np.linalg.eig(rate)
        inv_eigvecs = np.linalg.inv(eigvecs)
        self._setup_eigen(rate, eigvals, eigvecs, inv_eigvecs, device)
        self.base_rate, self.eigvals, self.eigvecs, self.inv_eigvecs = self._Rb, self._lam, self._U, self._Uinv

    def _integral_rate_scalar(self, t):
        return self.time_base * (self.time_exp ** t) - self.time_base

    def _rate_scalar(self, t):
        return self.time_base * math.log(self.time_exp) * (self.time_exp ** t)

    def rate(self, t):
        return self._scaled_rate(self._rate_scalar(t))

    def rate_mat(self, y, t):
        r = self.rate(t)
        bidx = utils.expand_dims(torch.arange(t.size(0), device=r.device), axis=tuple(range(1, y.dim())))
        return r[bidx, y.long()]

    def transition(self, t):
        return self._build_qt0(self._integral_rate_scalar(t))

    def transit_between(self, t1, t2):
        # the reference multiplies by eigvecs^T here, not inv_eigvecs (forward_model.py:298);
        # reproduced so ExactSampling-style callers see the same matrices.
        return self._build_qt0(self._integral_rate_scalar(t2) - self._integral_rate_scalar(t1), inverse=False)
