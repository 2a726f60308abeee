"""Thin tensor-level wrappers over the C ABI (include/ctdd.h): explicit tables in, tensors out.

The sampler / loss classes in lib/ build on these; tests call them directly to inject the oracle's q_{t|0}.
"""
from __future__ import annotations

import torch

from . import _native as nat


def build_qt0(U, Uinv, lam, d_int, normalize=True, clamp_below=1e-8, want_transpose=False):
    """Q_b = U diag(exp(lam*d_int[b])) Uinv -> (B,S,S) [and Q^T]."""
    B, S = d_int.shape[0], U.shape[0]
    Q = torch.empty((B, S, S), dtype=torch.float32, device=d_int.device)
    QT = torch.empty_like(Q) if want_transpose else None
    nat.check(nat.lib().ctdd_build_qt0(nat.ptr(U), nat.ptr(Uinv), nat.ptr(lam), nat.ptr(d_int.contiguous()), B, S,
                                       1 if normalize else 0, clamp_below, nat.ptr(Q), nat.ptr(QT), nat.stream()),
              "ctdd_build_qt0")
    return (Q, QT) if want_transpose else Q


def prep_tc_tables(Q, QT, Rb, eps, branch):
    """Per-time-point tables of the tcgen05 path; Q/QT are (T,S,S). Returns a (T, bytes) uint8 tensor or None."""
    T, S = Q.shape[0], Q.shape[-1]
    nbytes = int(nat.lib().ctdd_tc_tables_bytes(S))
    if nbytes <= 0:
        return None
    out = torch.empty((T, nbytes), dtype=torch.uint8, device=Q.device)
    nat.check(nat.lib().ctdd_prep_tc_tables(nat.ptr(Q), nat.ptr(QT), nat.ptr(Rb), T, S, float(eps), branch, nat.ptr(out),
                                            nat.stream()), "ctdd_prep_tc_tables")
    return out


def prep_tc_static(Rb):
    """Time-independent tables of the tcgen05 path (uint8 blob) or None when the path is unavailable for this S."""
    S = Rb.shape[-1]
    nbytes = int(nat.lib().ctdd_tc_static_bytes(S))
    if nbytes <= 0:
        return None
    out = torch.empty((nbytes,), dtype=torch.uint8, device=Rb.device)
    nat.check(nat.lib().ctdd_prep_tc_static(nat.ptr(Rb), S, nat.ptr(out), nat.stream()), "ctdd_prep_tc_static")
    return out


def reverse_step(mode, branch, logits, x_eval, Q, QT, Rb, RbT, beta, h, eps, *, N, D, S, x_base=None,
                 reject_multi=False, seed=0, offset=0, row_offset=0, impl=nat.IMPL_AUTO, tc_tables=None, tc_static=None,
                 workspace=None, stats=None, want_rr=False, want_ratio=False, logits_offset_elems=0,
                 batch_stride=None):
    """One fused reverse-rate evaluation (+ state update). Returns dict(x=..., rr=..., ratio=...)."""
    dev = x_eval.device
    if logits.dtype != torch.float32:
        logits = logits.float()
    logits = logits.contiguous()
    x_out = torch.empty((N, D), dtype=torch.int32, device=dev) if mode != nat.MODE_RATES_ONLY else None
    rr = torch.empty((N, D, S), dtype=torch.float32, device=dev) if want_rr else None
    ratio = torch.empty((N, D, S), dtype=torch.float32, device=dev) if want_ratio else None
    if workspace is None:
        ws = int(nat.lib().ctdd_step_workspace_bytes(N * D, S, impl))
        workspace = torch.empty((ws,), dtype=torch.uint8, device=dev) if ws > 0 else None
    p = nat.StepParams(
        mode=mode, branch=branch, impl=impl, N=N, D=D, S=S, row_offset=int(row_offset),
        logits=nat.ptr(logits) + 4 * int(logits_offset_elems), ld_logits=S,
        batch_stride_logits=(int(batch_stride) if batch_stride is not None else D * S),
        x_eval=nat.ptr(x_eval), x_base=nat.ptr(x_base),
        Q=nat.ptr(Q), QT=nat.ptr(QT), Rb=nat.ptr(Rb), RbT=nat.ptr(RbT), tc_tables=nat.ptr(tc_tables), tc_static=nat.ptr(tc_static),
        beta=float(beta), h=float(h), eps=float(eps), reject_multi=1 if reject_multi else 0,
        seed=int(seed), offset=int(offset),
        x_out=nat.ptr(x_out), rr_out=nat.ptr(rr), ratio_out=nat.ptr(ratio),
        stats_out=(stats.data_ptr() if stats is not None else None), workspace=nat.ptr(workspace))
    nat.check(nat.lib().ctdd_reverse_step(p, nat.stream()), "ctdd_reverse_step")
    return {"x": x_out, "rr": rr, "ratio": ratio}


def sample_categorical_shared(probs, rows, seed, offset=0, row_offset=0):
    x = torch.empty((rows,), dtype=torch.int32, device=probs.device)
    nat.check(nat.lib().ctdd_sample_categorical_shared(nat.ptr(probs.contiguous()), probs.shape[0], rows, row_offset,
                                                       int(seed), int(offset), nat.ptr(x), nat.stream()),
              "ctdd_sample_categorical_shared")
    return x


def noise_xt(Q, Rb, beta, x0, seed, offset=0, batch_offset=0, want_tilde=True):
    """x_t ~ Cat(Q[b, x0, :]) and the one-jump proposal x~ (lib/losses/losses.py:46-101). x0: (B,D) int32."""
    B, D = x0.shape
    S = Q.shape[-1]
    xt = torch.empty((B, D), dtype=torch.int32, device=x0.device)
    xtilde = torch.empty_like(xt) if want_tilde else None
    nat.check(nat.lib().ctdd_noise_xt(nat.ptr(Q.contiguous()), nat.ptr(Rb), nat.ptr(beta.contiguous()) if beta is not None else None,
                                      nat.ptr(x0.contiguous()), B, D, S, int(batch_offset), int(seed), int(offset),
                                      nat.ptr(xt), nat.ptr(xtilde), nat.stream()), "ctdd_noise_xt")
    return xt, xtilde


class _LossTerms(torch.autograd.Function):
    """Per-sample loss terms (out_a, out_b, out_c, out_d, out_nll), each (B,), differentiable w.r.t. `logits`
    through the fused backward kernel (include/ctdd.h: ctdd_loss_forward / ctdd_loss_backward)."""

    @staticmethod
    def forward(ctx, logits, kind, logit_branch, crm_type, Q, QT, Rb, beta, x0, xt, x_tilde, eps):
        B, D, S = logits.shape
        dev = logits.device
        logits = logits.contiguous().float()
        outs = torch.zeros((5, B), dtype=torch.float32, device=dev)
        nbytes = int(nat.lib().ctdd_loss_workspace_bytes(kind, B, S))
        ws = torch.empty((max(nbytes, 4),), dtype=torch.uint8, device=dev)
        p = nat.LossParams(kind=kind, logit_type=logit_branch, crm_type=crm_type, B=B, D=D, S=S,
                           logits=nat.ptr(logits), Q=nat.ptr(Q), QT=nat.ptr(QT), Rb=nat.ptr(Rb), beta=nat.ptr(beta),
                           x0=nat.ptr(x0), xt=nat.ptr(xt), x_tilde=nat.ptr(x_tilde), eps=float(eps),
                           out_a=outs[0].data_ptr(), out_b=outs[1].data_ptr(), out_c=outs[2].data_ptr(),
                           out_d=outs[3].data_ptr(), out_nll=outs[4].data_ptr(), workspace=nat.ptr(ws))
        nat.check(nat.lib().ctdd_loss_forward(p, nat.stream()), "ctdd_loss_forward")
        ctx.save_for_backward(logits, Q, QT, Rb, beta, x0, xt, x_tilde if x_tilde is not None else xt, ws)
        ctx.meta = (kind, logit_branch, crm_type, float(eps), x_tilde is not None)
        return outs[0], outs[1], outs[2], outs[3], outs[4]

    @staticmethod
    def backward(ctx, ga, gb, gc, gd, gn):
        logits, Q, QT, Rb, beta, x0, xt, x_tilde, ws = ctx.saved_tensors
        kind, logit_branch, crm_type, eps, has_tilde = ctx.meta
        B, D, S = logits.shape
        z = lambda g: (g if g is not None else torch.zeros((B,), device=logits.device)).contiguous().float()
        ga, gb, gd, gn = z(ga), z(gb), z(gd), z(gn)
        grad = torch.empty_like(logits)
        p = nat.LossParams(kind=kind, logit_type=logit_branch, crm_type=crm_type, B=B, D=D, S=S,
                           logits=nat.ptr(logits), Q=nat.ptr(Q), QT=nat.ptr(QT), Rb=nat.ptr(Rb), beta=nat.ptr(beta),
                           x0=nat.ptr(x0), xt=nat.ptr(xt), x_tilde=(nat.ptr(x_tilde) if has_tilde else None), eps=eps,
                           ga=nat.ptr(ga), gb=nat.ptr(gb), gd=nat.ptr(gd), gn=nat.ptr(gn), grad_logits=nat.ptr(grad),
                           workspace=nat.ptr(ws))
        nat.check(nat.lib().ctdd_loss_backward(p, nat.stream()), "ctdd_loss_backward")
        return (grad,) + (None,) * 11


def loss_terms(logits, kind, Q, QT, Rb, beta, x0, xt, x_tilde=None, eps=1e-9, logit_branch=nat.BRANCH_SDDM_REVERSE_PROB,
               crm_type=0):
    """(out_a, out_b, out_c, out_d, out_nll) per sample — see ctdd_loss_params in include/ctdd.h."""
    return _LossTerms.apply(logits, kind, logit_branch, crm_type, Q.contiguous(), QT.contiguous(), Rb, beta.contiguous().float(),
                            x0.contiguous(), xt.contiguous(), x_tilde.contiguous() if x_tilde is not None else None, eps)
