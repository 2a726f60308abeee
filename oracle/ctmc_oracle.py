"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch-CPU fp32 + numpy) of the reference's reverse-CTMC hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
Parity status: the reference ships no tests or golden vectors ("parity unpinned" by the reference itself);
this restatement is pinned against the reference's own code imported in the build container
(oracle/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).

Every function cites the reference file:line (relative to TAUnSDDM/) it follows.  Randomness is injected
through oracle/rng.py (Philox uniforms + inverse-CDF maps) exactly as the CUDA kernels consume it.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import rng

# ------------------------------------------------------------------------------------------------------
# forward process (lib/models/forward_model.py)


def gaussian_target_rate_matrix(S: int, rate_sigma: float, Q_sigma: float) -> np.ndarray:
    """forward_model.py:216-236 — banded Gaussian jump kernel with detailed balance w.r.t. N(S/2, Q_sigma)."""
    i = np.arange(S)[:, None]
    j = np.arange(S)[None, :]
    vals = np.exp(-np.arange(0, S, dtype=np.float64) ** 2 / (rate_sigma ** 2))
    rate = np.zeros((S, S))
    up = (i < S // 2) & (j > i) & (j < S - i)
    dn = (i > S // 2) & (j < i) & (j > -i + S - 1)
    rate[up] = vals[(j - i - 1)[up]]
    rate[dn] = vals[(i - j - 1)[dn]]
    # second sweep is sequential in the reference (row-major over (i, j), reading rate[j, i] as it goes)
    for a in range(S):
        for b in range(S):
            if rate[b, a] > 0.0:
                rate[a, b] = rate[b, a] * np.exp(
                    -((b + 1) ** 2 - (a + 1) ** 2 + S * (a + 1) - S * (b + 1)) / (2 * Q_sigma ** 2))
    rate = rate - np.diag(np.diag(rate))
    rate = rate - np.diag(np.sum(rate, axis=1))
    return rate


def uniform_rate_matrix(S: int, rate_const: float) -> np.ndarray:
    """forward_model.py:84-86."""
    rate = rate_const * np.ones((S, S))
    rate = rate - np.diag(np.diag(rate))
    return rate - np.diag(np.sum(rate, axis=1))


def birth_death_rate_matrix(S: int) -> np.ndarray:
    """forward_model.py:15-17."""
    r = np.diag(np.ones(S - 1), 1) + np.diag(np.ones(S - 1), -1)
    return r - np.diag(np.sum(r, axis=1))


class ForwardProcess:
    """rate(t), transition(t)=q_{t|0}, rate_mat, transit_between for the four rate families.

    kind: 'gaussian' (GaussianTargetRate :207-306), 'uniform' (UniformRate :78-129),
          'uniform_variant' (UniformVariantRate :132-204), 'birth_death' (BirthDeathForwardBase :9-75).
    """

    def __init__(self, kind: str, S: int, **kw):
        self.kind, self.S = kind, S
        if kind == "gaussian":
            self.rate_sigma, self.Q_sigma = kw["rate_sigma"], kw["Q_sigma"]
            self.time_exp, self.time_base = kw["time_exp"], kw["time_base"]
            rate = gaussian_target_rate_matrix(S, self.rate_sigma, self.Q_sigma)
            eigvals, eigvecs = np.linalg.eig(rate)
            inv = np.linalg.inv(eigvecs)
        elif kind in ("uniform", "uniform_variant"):
            self.rate_const = kw["rate_const"]
            self.t_func = kw.get("t_func")
            self.time_exp, self.time_base = kw.get("time_exp"), kw.get("time_base")
            rate = uniform_rate_matrix(S, self.rate_const)
            eigvals, eigvecs = np.linalg.eigh(rate)
            inv = eigvecs.T
        elif kind == "birth_death":
            self.sigma_min, self.sigma_max = kw["sigma_min"], kw["sigma_max"]
            rate = birth_death_rate_matrix(S)
            eigvals, eigvecs = np.linalg.eigh(rate)
            inv = eigvecs.T
        else:
            raise ValueError(kind)
        self.base_rate = torch.from_numpy(np.real(rate)).float()
        self.eigvals = torch.from_numpy(np.real(eigvals)).float()
        self.eigvecs = torch.from_numpy(np.real(eigvecs)).float()
        self.inv_eigvecs = torch.from_numpy(np.real(inv)).float()

    # --- time warps -----------------------------------------------------------------------------------
    def int_beta(self, t: torch.Tensor) -> torch.Tensor:
        if self.kind == "gaussian" or (self.kind == "uniform_variant" and self.t_func == "log"):
            return self.time_base * (self.time_exp ** t) - self.time_base  # :246-247, :150
        if self.kind == "uniform":
            return t  # :114 exp(eigvals * t)
        if self.kind == "birth_death":
            return 0.5 * self.sigma_min ** 2 * (self.sigma_max / self.sigma_min) ** (2 * t) - 0.5 * self.sigma_min ** 2
        if self.t_func == "log_sqr":
            return torch.log(t ** 2 + 1)  # :146
        if self.t_func == "sqrt_cos":
            return -torch.sqrt(torch.cos(torch.pi / 2 * t))  # :148
        raise ValueError(f"Unknown t_func {self.t_func}")

    def beta(self, t: torch.Tensor) -> torch.Tensor:
        if self.kind == "gaussian" or (self.kind == "uniform_variant" and self.t_func == "log"):
            return self.time_base * math.log(self.time_exp) * (self.time_exp ** t)  # :249-250, :162
        if self.kind == "uniform":
            return torch.ones_like(t)
        if self.kind == "birth_death":
            return self.sigma_min ** 2 * (self.sigma_max / self.sigma_min) ** (2 * t) * math.log(self.sigma_max / self.sigma_min)
        if self.t_func == "log_sqr":
            return 2 * t / (t ** 2 + 1)  # :156
        if self.t_func == "sqrt_cos":
            tt = torch.pi / 2 * t
            return torch.pi / 4.0 * (torch.sin(tt) / torch.sqrt(torch.cos(tt)))  # :158-160
        raise ValueError(f"Unknown t_func {self.t_func}")

    def rate(self, t: torch.Tensor) -> torch.Tensor:
        """R_t = beta(t) R_b, (B,S,S)  (:252-257, :95-101, :166-172, :43-49)."""
        return self.base_rate.view(1, self.S, self.S) * self.beta(t).view(-1, 1, 1)

    def rate_mat(self, y: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """rows R_t[b, y] (:103-105, :174-178, :259-263)."""
        r = self.rate(t)
        b = torch.arange(t.shape[0]).view(-1, *([1] * (y.dim() - 1)))
        return r[b, y.long()]

    def _expm(self, d_int: torch.Tensor, use_inverse: bool) -> torch.Tensor:
        B, S = d_int.shape[0], self.S
        right = self.inv_eigvecs if use_inverse else self.eigvecs.T
        return (self.eigvecs.view(1, S, S) @ torch.diag_embed(torch.exp(d_int.view(B, 1) * self.eigvals.view(1, S)))
                @ right.reshape(1, S, S))

    def transition(self, t: torch.Tensor) -> torch.Tensor:
        """q_{t|0}, (B,S,S), Q[b,k,s] = P(x_t = s | x_0 = k)  (:265-287, :108-126, :202-204, :51-75)."""
        tr = self._expm(self.int_beta(t) - (self.int_beta(torch.zeros_like(t)) if self.kind == "uniform_variant" else 0.0),
                        use_inverse=True)
        if self.kind != "uniform":  # UniformRate.transition does not renormalise (:108-126)
            tr = tr / torch.sum(tr, dim=-1, keepdim=True)
        tr = torch.where(tr < 1e-8, torch.zeros_like(tr), tr)
        return tr

    def transit_between(self, t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
        """q_{t2|t1} (:128-129, :180-200, :289-306; the Gaussian family multiplies by eigvecs^T, quirk B.7)."""
        if self.kind == "uniform":
            return self.transition(t2 - t1)
        tr = self._expm(self.int_beta(t2) - self.int_beta(t1), use_inverse=(self.kind != "gaussian"))
        tr = tr / torch.sum(tr, dim=-1, keepdim=True)
        return torch.where(tr < 1e-8, torch.zeros_like(tr), tr)


# ------------------------------------------------------------------------------------------------------
# reverse rates (lib/sampling/sampling.py:31-78, lib/models/model_utils.py:30-60)

TAULDR_LOSSES = ("CTElbo", "NLL", "CTElboLambda")


def logprob_with_logits(logits: torch.Tensor, xt: torch.Tensor, Q: Optional[torch.Tensor], logit_type: str):
    """model_utils.py:30-60 -> (ll_all (B,D,S), ll_xt (B,D)). Q is (B,S,S)."""
    S = logits.shape[-1]
    if logit_type == "direct":
        ll = F.log_softmax(logits, dim=-1)
    elif logit_type == "reverse_prob":
        ll = torch.log(F.softmax(logits, dim=-1) @ Q + 1e-35)
    elif logit_type == "reverse_logscale":
        lq = torch.where(Q <= 1e-35, torch.full_like(Q, -1e9), torch.log(Q))
        ll = torch.logsumexp(F.log_softmax(logits, dim=-1).unsqueeze(-1) + lq.unsqueeze(1), dim=-2)
    else:
        raise ValueError("Unknown logit_type: %s" % logit_type)
    ll_xt = torch.sum(ll * F.one_hot(xt.long(), S), dim=-1)
    return ll, ll_xt


def reverse_rates(logits: torch.Tensor, x: torch.Tensor, Q: torch.Tensor, R: torch.Tensor, loss_name: str,
                  logit_type: str = "reverse_prob", eps: float = 1e-9):
    """sampling.py:31-78. logits (N,D,S), x (N,D), Q/R (N,S,S) or (1,S,S). Returns (rr, ratio), s==x NOT zeroed."""
    N, D, S = logits.shape
    Q = Q.expand(N, S, S)
    R = R.expand(N, S, S)
    n = torch.arange(N).view(N, 1)
    xl = x.long()
    if loss_name in TAULDR_LOSSES:
        p0t = F.softmax(logits, dim=2)
        den = Q.transpose(1, 2)[n, xl] + eps          # Q[n, :, x]  (column x)
        fwd = R.transpose(1, 2)[n, xl]                # R[n, :, x]  (column x: rate s -> x)
        ratio = (p0t / den) @ Q
        return fwd * ratio, ratio
    ll, llx = logprob_with_logits(logits, x, Q, logit_type)
    ratio = torch.exp(ll - llx.unsqueeze(-1))
    return ratio * R[n, xl], ratio                    # row x


# ------------------------------------------------------------------------------------------------------
# state updates with injected uniforms


# "map": jump counts through the shared uniform -> sample map (rng.poisson_rows; bit-comparable with the kernels);
# "torch": through torch.poisson exactly as the reference draws them (free-running distributional checks)
POISSON_LAW = "map"
_TORCH_GEN = torch.Generator().manual_seed(0)


def set_poisson_law(law: str, seed: int = 0):
    global POISSON_LAW
    assert law in ("map", "torch")
    POISSON_LAW = law
    _TORCH_GEN.manual_seed(seed)


def _zero_at(rr: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return rr * (1 - F.one_hot(x.long(), rr.shape[-1])).to(rr.dtype)


def tau_leap_update(rates_z: torch.Tensor, x_eval: torch.Tensor, x_base: torch.Tensor, h: float, S: int,
                    reject_multi: bool, offset: int, seed: int, row_offset: int = 0):
    """sampling.py:127-160 (and :481-503 for midpoint stage 2). rates_z: (N,D,S) with s==x_eval zeroed."""
    N, D, _ = rates_z.shape
    lam = (rates_z.detach() * h).numpy().reshape(N * D, S)
    if POISSON_LAW == "torch":     # the reference's own draw (torch.poisson, sampling.py:131): distributional checks only
        kc = torch.poisson(torch.from_numpy(np.ascontiguousarray(lam)), generator=_TORCH_GEN).numpy().astype(np.int64)
    else:
        kc, _ = rng.poisson_rows(lam, row_offset, offset, seed)
    diff = np.arange(S, dtype=np.int64)[None, :] - x_eval.numpy().reshape(-1, 1).astype(np.int64)
    jump = (kc * diff).sum(axis=1)
    cnt = kc.sum(axis=1)
    stats = {"rows_jumped": int((cnt > 0).sum()), "rows_multi": int((cnt > 1).sum())}
    if reject_multi:
        jump = np.where(cnt > 1, 0, jump)
    xb = x_base.numpy().reshape(-1).astype(np.int64)
    xn = np.clip(xb + jump, 0, S - 1)
    stats["nonzero_jump"] = int((jump != 0).sum())
    stats["changed_base"] = int((xn != xb).sum())
    stats["changed_eval"] = int((xn != x_eval.numpy().reshape(-1)).sum())
    return torch.from_numpy(xn.reshape(N, D)), stats


def tau_leap_margin(rates_z: torch.Tensor, h: float, offset: int, seed: int, row_offset: int = 0) -> np.ndarray:
    """(N*D,) tie margin of every row of tau_leap_update on the same inputs (rng.poisson_rows_margin)."""
    N, D, S = rates_z.shape
    lam = (rates_z.detach() * h).numpy().reshape(N * D, S)
    return rng.poisson_rows_margin(lam, row_offset, offset, seed)


def euler_margin(rates_z: torch.Tensor, x: torch.Tensor, h: float, S: int, offset: int, seed: int, row_offset: int = 0):
    """(N*D,) tie margin of every row of euler_update on the same inputs."""
    N, D, _ = rates_z.shape
    r = rates_z.detach().numpy().reshape(N * D, S).astype(np.float32)
    tot = np.cumsum(r, axis=1, dtype=np.float32)[:, -1]
    diag = np.maximum(np.float32(0.0), (np.float32(1.0) - np.float32(h) * tot).astype(np.float32))
    P = (r * np.float32(h)).astype(np.float32)
    P[np.arange(N * D), x.numpy().reshape(-1).astype(np.int64)] = diag
    return rng.inv_cdf_margin(P, rng.row_units(N * D, row_offset, offset, rng.STREAM_ROW, seed))


def midpoint_drift_margin(rates_z: torch.Tensor, x: torch.Tensor, h: float, S: int) -> np.ndarray:
    """(N*D,) distance of 0.5*h*sum_s rr_s*(s-x) to the nearest rounding boundary (k + 1/2), relative to sum_s rr_s*|s-x|*h/2."""
    diff = (torch.arange(S).view(1, 1, S) - x.long().unsqueeze(-1)).to(torch.float64)
    val = (0.5 * h * torch.sum(rates_z.double() * diff, dim=-1)).numpy().reshape(-1)
    scale = (0.5 * h * torch.sum(rates_z.double() * diff.abs(), dim=-1)).numpy().reshape(-1)
    frac = np.abs(val - np.floor(val) - 0.5)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(scale > 0, frac / np.where(scale > 0, scale, 1.0), np.inf)


def euler_update(rates_z: torch.Tensor, x: torch.Tensor, h: float, S: int, offset: int, seed: int, row_offset: int = 0):
    """sampling.py:278-293: P = rr*h off-diagonal, max(0, 1 - h*sum) on the diagonal; categorical draw."""
    N, D, _ = rates_z.shape
    r = rates_z.detach().numpy().reshape(N * D, S).astype(np.float32)
    tot = np.cumsum(r, axis=1, dtype=np.float32)[:, -1]
    diag = np.maximum(np.float32(0.0), (np.float32(1.0) - np.float32(h) * tot).astype(np.float32))
    P = (r * np.float32(h)).astype(np.float32)
    xi = x.numpy().reshape(-1).astype(np.int64)
    P[np.arange(N * D), xi] = diag
    v = rng.row_units(N * D, row_offset, offset, rng.STREAM_ROW, seed)
    xn = rng.inv_cdf(P, v)
    return torch.from_numpy(xn.reshape(N, D)), {"changed_base": int((xn != xi).sum())}


def midpoint_drift(rates_z: torch.Tensor, x: torch.Tensor, h: float, S: int):
    """sampling.py:433-453: x' = clip(x + int(round_half_even(0.5*h*sum_s rr_s*(s-x))))."""
    diff = (torch.arange(S).view(1, 1, S) - x.long().unsqueeze(-1)).to(torch.float32)
    change = torch.round(0.5 * h * torch.sum(rates_z * diff, dim=-1)).to(torch.int64)
    return torch.clamp(x.long() + change, 0, S - 1)


# ------------------------------------------------------------------------------------------------------
# samplers

ModelFn = Callable[[torch.Tensor, torch.Tensor], torch.Tensor]  # (x (N,D) int64, t (N,) fp32) -> logits (N,D,S)


def initial_probs(S: int, initial_dist: str, std: Optional[float]) -> np.ndarray:
    """sampling.py:14-28 target distribution as fp32 weights."""
    if initial_dist == "uniform":
        return np.full(S, 1.0 / S, dtype=np.float32)
    if initial_dist == "gaussian":
        t = np.exp(-((np.arange(1, S + 1) - S // 2) ** 2) / (2 * std ** 2))
        return (t / np.sum(t)).astype(np.float32)
    raise NotImplementedError("Unrecognized initial dist " + initial_dist)


def initial_samples(N: int, D: int, S: int, initial_dist: str, std, seed: int, row_offset: int = 0) -> torch.Tensor:
    p = initial_probs(S, initial_dist, std)
    v = rng.row_units(N * D, row_offset, 0, rng.STREAM_INIT, seed)
    x = rng.inv_cdf(np.broadcast_to(p, (N * D, S)), v)
    return torch.from_numpy(x.reshape(N, D))


def _rates_at(fp: ForwardProcess, model: ModelFn, x, t: float, N, loss_name, logit_type, eps, corrector=False):
    t_ones = t * torch.ones((N,))
    Q = fp.transition(t_ones[:1])
    R = fp.rate(t_ones[:1])
    logits = model(x.long(), t_ones)
    rr, _ = reverse_rates(logits, x, Q, R, loss_name, logit_type, eps)
    rz = _zero_at(rr, x)
    if corrector:  # sampling.py:183-198: R_t[x, :] + rr, diagonal zeroed
        rz = _zero_at(R.expand(N, -1, -1)[torch.arange(N).view(N, 1), x.long()] + rz, x)
    return rz


def _final_argmax(model: ModelFn, x, min_t: float, N: int):
    return torch.max(F.softmax(model(x.long(), min_t * torch.ones((N,))), dim=2), dim=2)[1]


@torch.no_grad()
def sample_taul(fp, model, N, D, S, *, max_t, min_t, num_steps, initial_dist, init_std, is_ordinal, loss_name,
                logit_type="reverse_prob", eps=1e-9, corrector_entry_time=0.0, num_corrector_steps=0, seed=0):
    """TauL.sample  sampling.py:98-234 -> (x (N,D) int64 ndarray, change_dim list)."""
    x = initial_samples(N, D, S, initial_dist, init_std, seed)
    ts = np.concatenate((np.linspace(max_t, min_t, num_steps), np.array([0])))
    change_dim, call = [], 0
    for idx, t in enumerate(ts[0:-1]):
        h = ts[idx] - ts[idx + 1]
        rz = _rates_at(fp, model, x, t, N, loss_name, logit_type, eps)
        x_new, st = tau_leap_update(rz, x, x, h, S, not is_ordinal, call, seed)
        call += 1
        change_dim.append(st["changed_base"] / N)
        x = x_new
        if t <= corrector_entry_time:
            for _ in range(num_corrector_steps):
                rz = _rates_at(fp, model, x, t, N, loss_name, logit_type, eps, corrector=True)
                x, _ = tau_leap_update(rz, x, x, h, S, not is_ordinal, call, seed)
                call += 1
    if loss_name in ("CTElbo", "NLL"):
        x = _final_argmax(model, x, min_t, N)
    return x.numpy().astype(int), change_dim


@torch.no_grad()
def sample_lbjf(fp, model, N, D, S, *, max_t, min_t, num_steps, initial_dist, init_std, loss_name,
                logit_type="reverse_prob", eps=1e-9, corrector_entry_time=0.0, num_corrector_steps=0, seed=0):
    """LBJF.sample  sampling.py:253-356."""
    x = initial_samples(N, D, S, initial_dist, init_std, seed)
    ts = np.concatenate((np.linspace(max_t, min_t, num_steps), np.array([0])))
    change_dim, call = [], 0
    for idx, t in enumerate(ts[0:-1]):
        h = ts[idx] - ts[idx + 1]
        rz = _rates_at(fp, model, x, t, N, loss_name, logit_type, eps)
        x_new, st = euler_update(rz, x, h, S, call, seed)
        call += 1
        change_dim.append(st["changed_base"] / N)
        if t <= corrector_entry_time:
            for _ in range(num_corrector_steps):
                rz = _rates_at(fp, model, x_new, t, N, loss_name, logit_type, eps, corrector=True)
                x_new, _ = euler_update(rz, x_new, h, S, call, seed)
                call += 1
        x = x_new
    if loss_name == "CTElbo":
        x = _final_argmax(model, x, min_t, N)
    return x.numpy().astype(int), change_dim


@torch.no_grad()
def sample_midpoint(fp, model, N, D, S, *, max_t, min_t, num_steps, initial_dist, init_std, is_ordinal, loss_name,
                    logit_type="reverse_prob", eps=1e-9, seed=0):
    """MidPointTauL.sample  sampling.py:390-526 (state_change table generalised to s - x)."""
    x = initial_samples(N, D, S, initial_dist, init_std, seed)
    t = max_t
    h = (max_t - min_t) / num_steps
    change_jump, change_dim, change_dim_first, change_1to2, call = [], [], [], [], 0
    while t - 0.5 * h > min_t:
        rz = _rates_at(fp, model, x, t, N, loss_name, logit_type, eps)
        x_prime = midpoint_drift(rz, x, h, S)
        change_dim_first.append(float((x != x_prime).sum()) / (N * D))
        # stage 2 evaluates at t - h/2 with a fp32 time vector: t_ones - 0.5*h (sampling.py:416)
        t05 = float((torch.tensor(t, dtype=torch.float32) * torch.ones(1) - 0.5 * h)[0])
        rz2 = _rates_at(fp, model, x_prime, t05, N, loss_name, logit_type, eps)
        x_new, st = tau_leap_update(rz2, x_prime, x, h, S, not is_ordinal, call, seed)
        call += 1
        if is_ordinal:
            change_jump.append(st["rows_multi"] / st["rows_jumped"] if st["rows_jumped"] else float("nan"))
        change_dim.append(st["nonzero_jump"] / (N * D))
        change_1to2.append(st["changed_eval"] / (N * D))
        x = x_new
        t = t - h
    if loss_name == "CTElbo":
        x = _final_argmax(model, x, min_t, N)
    return x.numpy().astype(int), change_jump, change_dim, change_dim_first, change_1to2


@torch.no_grad()
def sample_pctaul(fp, model, N, D, S, *, min_t, num_steps, initial_dist, eps=1e-9, corrector_entry_time=0.0,
                  num_corrector_steps=0, corrector_step_size_multiplier=1.0, seed=0):
    """PCTauL.sample  sampling.py:534-646: tauLDR rates regardless of loss.name, no rejection, always argmax."""
    x = initial_samples(N, D, S, initial_dist, 200, seed)  # std hard-coded :548
    h0 = 1.0 / num_steps
    ts = np.linspace(1.0, min_t + h0, num_steps)
    call = 0
    for idx, t in enumerate(ts[0:-1]):
        h = ts[idx] - ts[idx + 1]
        rz = _rates_at(fp, model, x, t, N, "CTElbo", "direct", eps)
        x, _ = tau_leap_update(rz, x, x, h, S, False, call, seed)
        call += 1
        if t <= corrector_entry_time:
            for _ in range(num_corrector_steps):
                rz = _rates_at(fp, model, x, t - h, N, "CTElbo", "direct", eps, corrector=True)
                x, _ = tau_leap_update(rz, x, x, corrector_step_size_multiplier * h, S, False, call, seed)
                call += 1
    return _final_argmax(model, x, min_t, N).numpy().astype(int)


@torch.no_grad()
def sample_conditional_taul(fp, model, N, total_D, S, conditioner: torch.Tensor, *, condition_dim, min_t, num_steps,
                            initial_dist, init_std, eps=1e-9, seed=0):
    """ConditionalTauLeaping.sample  sampling.py:654-758 (rejection mask is overwritten there: no rejection)."""
    sample_D = total_D - condition_dim
    x = initial_samples(N, sample_D, S, initial_dist, init_std, seed)
    ts = np.concatenate((np.linspace(1.0, min_t, num_steps), np.array([0])))

    def sliced(xx, tt):
        return model(torch.concat((conditioner.long(), xx), dim=1), tt)[:, condition_dim:, :]

    call = 0
    for idx, t in enumerate(ts[0:-1]):
        h = ts[idx] - ts[idx + 1]
        rz = _rates_at(fp, sliced, x, t, N, "CTElbo", "direct", eps)
        x, _ = tau_leap_update(rz, x, x, h, S, False, call, seed)
        call += 1
    x0 = _final_argmax(sliced, x, min_t, N)
    return torch.concat((conditioner.long(), x0), dim=1).numpy().astype(int)


@torch.no_grad()
def sample_conditional_pctaul(fp, model, N, total_D, S, conditioner: torch.Tensor, *, condition_dim, min_t, num_steps,
                              initial_dist, init_std, reject=False, eps=1e-9, corrector_entry_time=0.0,
                              num_corrector_steps=0, corrector_step_size_multiplier=1.0, seed=0):
    """ConditionalPCTauLeaping.sample  sampling.py:766-905."""
    sample_D = total_D - condition_dim
    x = initial_samples(N, sample_D, S, initial_dist, init_std, seed)
    h0 = 1.0 / num_steps
    ts = np.linspace(1.0, min_t + h0, num_steps)

    def sliced(xx, tt):
        return model(torch.concat((conditioner.long(), xx), dim=1), tt)[:, condition_dim:, :]

    call = 0
    for idx, t in enumerate(ts[0:-1]):
        h = ts[idx] - ts[idx + 1]
        rz = _rates_at(fp, sliced, x, t, N, "CTElbo", "direct", eps)
        x, _ = tau_leap_update(rz, x, x, h, S, reject, call, seed)
        call += 1
        if t <= corrector_entry_time:
            for _ in range(num_corrector_steps):
                rz = _rates_at(fp, sliced, x, t - h, N, "CTElbo", "direct", eps, corrector=True)
                x, _ = tau_leap_update(rz, x, x, corrector_step_size_multiplier * h, S, reject, call, seed)
                call += 1
    x0 = _final_argmax(sliced, x, min_t, N)
    return torch.concat((conditioner.long(), x0), dim=1).numpy().astype(int)


def sample_exact(fp, model, N, D, S, *, max_t, min_t, num_steps, initial_dist, init_std, seed):
    """ExactSampling.sample (sampling.py:994-1061): per step x ~ Cat(logits = logsumexp_k(log p0t[k] +
    log(q_{t-h|0}[k, s'] * q_{t|t-h}[s', x]))), one categorical per row (per-row stream, call offset = step)."""
    xt = initial_samples(N, D, S, initial_dist, init_std, seed)
    ts = np.concatenate((np.linspace(max_t, min_t, num_steps), np.array([0])))
    change = []
    n = torch.arange(N).view(N, 1)
    for idx, t in enumerate(ts[0:-1]):
        h = ts[idx] - ts[idx + 1]
        t_ones = t * torch.ones((N,))
        p0t = F.softmax(model(xt, t_ones), dim=2)
        t_eps = t - h
        q_teps_0 = fp.transition(t_eps * torch.ones((1,)))[0]                        # (S,S)
        q_t_teps = fp.transit_between(t_eps * torch.ones((1,)), t * torch.ones((1,)))[0]  # (S,S): [s', x]
        col = q_t_teps.t()[xt.long()]                                                 # (N,D,S): q_{t|t-h}[s', x]
        w = (p0t @ q_teps_0) * col
        v = rng.row_units(N * D, 0, idx, rng.STREAM_ROW, seed)
        x_new = torch.from_numpy(rng.inv_cdf(w.detach().numpy().reshape(N * D, S).astype(np.float32), v).reshape(N, D))
        change.append(float((x_new != xt).sum()) / (N * D))
        xt = x_new
    return xt.numpy().astype(int), change

