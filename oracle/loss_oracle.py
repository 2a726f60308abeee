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


# ------------------------------------------------------------------------------------------------------
# loss terms (torch-CPU, differentiable)

import torch.nn.functional as F  # noqa: E402

from . import ctmc_oracle as oc  # noqa: E402


def ctelbo_terms(p_reg, p_sig, Q, R, x0, reg_x, x_tilde, eps):
    """losses.py:121-282 -> (reg (B,), outer (B,), norm (B,)). p_*: (B,D,S) softmax outputs; Q,R: (B,S,S)."""
    B, D, S = p_reg.shape
    bi = torch.arange(B).view(B, 1)
    QT, RT = Q.transpose(1, 2), R.transpose(1, 2)
    mask_reg = 1.0 - F.one_hot(reg_x.long(), S).float()
    reg_tmp = (mask_reg * RT[bi, reg_x.long()]) @ QT                     # :148
    reg = torch.sum((p_reg / (QT[bi, reg_x.long()] + eps)) * reg_tmp, dim=(1, 2))   # :153
    xt = x_tilde.long()
    inner = torch.log((p_sig / (QT[bi, xt] + eps)) @ Q + eps)            # :179-181
    mask = 1.0 - F.one_hot(xt, S).float()
    w = mask * RT[bi, xt] * (Q[bi, x0.long()] / (Q[bi, x0.long(), xt] + eps).unsqueeze(-1))
    outer = torch.sum(w * inner, dim=(1, 2))                             # :215-221
    z = -torch.diagonal(R, dim1=1, dim2=2)                               # :225-229
    zx = z[bi, xt]
    Z = zx.sum(1).view(B, 1, 1) - zx.unsqueeze(-1) + z.view(B, 1, S)      # :242-246
    norm = torch.sum(w / Z, dim=(1, 2))                                  # :272-276
    return reg, outer, norm


def sddm_terms(ll_all, ll_x, Q, R, x0, x_tilde, eps):
    """losses.py:1375-1486 -> (reg, outer, norm)."""
    B, D, S = ll_all.shape
    bi = torch.arange(B).view(B, 1)
    RT = R.transpose(1, 2)
    xt = x_tilde.long()
    mask = 1.0 - F.one_hot(xt, S).float()
    L = ll_all - ll_x.unsqueeze(-1)
    reg = torch.sum(torch.exp(L) * mask * RT[bi, xt], dim=(1, 2))
    w = mask * RT[bi, xt] * (Q[bi, x0.long()] / (Q[bi, x0.long(), xt] + eps).unsqueeze(-1))
    outer = torch.sum(w * L, dim=(1, 2))
    z = -torch.diagonal(R, dim1=1, dim2=2)
    zx = z[bi, xt]
    Z = zx.sum(1).view(B, 1, 1) - zx.unsqueeze(-1) + z.view(B, 1, S)
    return reg, outer, torch.sum(w / Z, dim=(1, 2))


def log1mexp(x):
    """lib/utils/utils.py:86-91."""
    x = -torch.abs(x)
    return torch.where(x > -0.693, torch.log(-torch.expm1(x)), torch.log1p(-torch.exp(x)))


def crm_rows(loss_type, ll_all, ll_x, Q, xt, S):
    """losses.py:794-836 per-(b,d) loss."""
    if loss_type == "rm":
        return -ll_x
    if loss_type == "mle":
        return -((S - 1) * ll_x + torch.sum(log1mexp(ll_all), dim=-1) - log1mexp(ll_x))
    if loss_type == "elbo":
        B = ll_all.shape[0]
        bi = torch.arange(B).view(B, 1)
        oh = F.one_hot(xt.long(), S).float()
        e = torch.exp(ll_all - ll_x.unsqueeze(-1))
        first = torch.sum(e * Q.transpose(1, 2)[bi, xt.long()] * (1 - oh), dim=-1)
        second = torch.sum((ll_x.unsqueeze(-1) - ll_all) * Q[bi, xt.long()] * (1 - oh), dim=-1)
        return first - second
    raise ValueError("Unknown loss_type: %s" % loss_type)


def loss_value(name, fp, model, x0, ts, *, seed, offset=0, eps=1e-9, nll_weight=0.001, logit_type="reverse_prob",
               loss_type="rm", ce_coeff=0.0, one_forward_pass=True, n_iter=0, n_iters=1, condition_dim=0, label=None):
    """Scalar loss of the named reference class for a given time draw `ts` and injected noising uniforms.

    model(x (B,D) int64, t (B,)) -> logits (B,D,S), differentiable. Mirrors calc_loss of CTElbo (:22-287),
    NLL (:1514-1778), CTElboLambda (:1794-2058), CondCTElbo (:558-781), CatRM (:838-890), CatRMNLL (:1190-1242),
    ScoreElbo (:1255-1500), SDDMElbo (:300-544), NLLOriginal (:1059-1103)."""
    cond = None
    if name == "CondCTElbo":
        cond, x0 = x0[:, :condition_dim], x0[:, condition_dim:]
    B, D = x0.shape
    S = fp.S
    Q, R = fp.transition(ts), fp.rate(ts)
    xt, xtil = noise_xt(Q, R, x0, seed, offset)
    ce = lambda lg: F.cross_entropy(lg.permute(0, 2, 1), x0.long())
    if name in ("CTElbo", "NLL", "CTElboLambda"):
        logits = model(xt, ts)
        p = F.softmax(logits, dim=2)
        if one_forward_pass:
            reg, outer, norm = ctelbo_terms(p, p, Q, R, x0, xtil, xtil, eps)
        else:
            p_sig = F.softmax(model(xtil, ts), dim=2)
            reg, outer, norm = ctelbo_terms(p, p_sig, Q, R, x0, xt, xtil, eps)
        neg_elbo = torch.mean(-outer / norm) + torch.mean(reg)
        nll = ce(logits)
        if name == "CTElbo":
            return neg_elbo + nll_weight * nll
        if name == "NLL":
            return nll
        w = n_iter / n_iters
        return w * neg_elbo + (1 - w) * nll
    if name == "CondCTElbo":
        sl = lambda x: model(torch.concat((cond.long(), x), dim=1), ts)[:, condition_dim:, :]
        if one_forward_pass:
            logits = sl(xtil)
            p = F.softmax(logits, dim=2)
            reg, outer, norm = ctelbo_terms(p, p, Q, R, x0, xtil, xtil, eps)
        else:
            # losses.py:660-666: the second pass OVERWRITES x_logits, so the cross-entropy (:777-779) sees the logits of x~
            logits_reg = sl(xt)
            logits = sl(xtil)
            reg, outer, norm = ctelbo_terms(F.softmax(logits_reg, dim=2), F.softmax(logits, dim=2), Q, R, x0, xt, xtil, eps)
        return torch.mean(-outer / norm) + torch.mean(reg) + nll_weight * ce(logits)
    if name in ("ScoreElbo", "SDDMElbo"):
        logits = model(xtil, ts)
        ll_all, ll_x = oc.logprob_with_logits(logits, xtil, Q, logit_type)
        reg, outer, norm = sddm_terms(ll_all, ll_x, Q, R, x0, xtil, eps)
        neg_elbo = torch.mean(-outer / norm) + torch.mean(reg)
        if name == "ScoreElbo":
            return neg_elbo + nll_weight * torch.sum(-ll_x) / B
        return neg_elbo + nll_weight * ce(logits)
    if name in ("CatRM", "CatRMNLL"):
        logits = model(xt, ts)
        ll_all, ll_x = oc.logprob_with_logits(logits, xt, Q, logit_type)
        loss = torch.sum(crm_rows(loss_type, ll_all, ll_x, Q, xt, S) * (1 - ce_coeff)) / B
        return loss if name == "CatRM" else loss + nll_weight * ce(logits)
    if name == "NLLOriginal":
        return ce(model(xt, ts, label))
    raise KeyError(name)
