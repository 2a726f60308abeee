"""Diagnostic: rows where the tensor-core and the CUDA-core loss contraction disagree (not part of the product)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctdd_b200 import _native as nat, make_config, ops
from ctdd_b200.lib.models import forward_model as fm
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
S, D, B = 256, int(os.environ.get("D", 96)), int(os.environ.get("B", 5))
cfg = make_config(data=dict(S=S), model=dict(rate_sigma=6.0, Q_sigma=512.0, time_exp=100.0, time_base=3.0), device="cuda")
model = fm.GaussianTargetRate(cfg, "cuda")
g = torch.Generator(device="cuda").manual_seed(1)
ts = torch.rand(B, device="cuda", generator=g) * 0.9 + 0.05
Q, QT = model._build_qt0(model._transition_delta(ts), inverse=True, want_transpose=True)
beta = model._rate_scalar(ts).float().contiguous()
Rb, _ = model.base_rate_tables(Q.device)
x0 = torch.randint(0, S, (B, D), device="cuda", generator=g, dtype=torch.int32)
xt, xtil = ops.noise_xt(Q, Rb, beta, x0, 5, 0)
logits = torch.randn((B, D, S), device="cuda", generator=g) - (torch.arange(S, device="cuda").view(1, 1, S) - x0.unsqueeze(-1)).float() ** 2 / 128.0
p = torch.softmax(logits.double(), -1)
u64 = torch.einsum("bdk,bks->bds", p, Q.double())
ux64 = torch.gather(u64, 2, xt.long().unsqueeze(-1)).squeeze(-1)
print("ts", ts.tolist())
outs = {}
for tc in (False, True):
    ops._LossTerms.use_tc = tc
    lg = logits.clone().requires_grad_(True)
    o = ops.loss_terms(lg, nat.LOSS_CRM, Q=Q, QT=QT, Rb=Rb, beta=beta, x0=x0, eps=1e-9, xt=xt, crm_type=0)
    outs[tc] = [t.detach().double().cpu() for t in o]
    if tc:
        scr = o[0].grad_fn.scr.view(torch.float32)
        n = B * D * S
        bufU = scr[n:2 * n].view(B, D, S).double()
        uxtc = torch.gather(bufU, 2, xt.long().unsqueeze(-1)).squeeze(-1)
        rel = ((uxtc - ux64).abs() / ux64.clamp_min(1e-300))
        print("tc u_x vs fp64: max rel", rel.max().item(), "rows with rel > 1e-4:", int((rel > 1e-4).sum()))
        idx = torch.nonzero(rel > 1e-4)[:10]
        for b, d in idx.tolist():
            print("  b", b, "d", d, "x0", int(x0[b, d]), "xt", int(xt[b, d]), "ux64", ux64[b, d].item(), "tc", uxtc[b, d].item())
        allrel = ((bufU - u64).abs() / u64.clamp_min(1e-300))
        print("all entries: max rel", allrel.max().item(), " frac > 1e-4:", (allrel > 1e-4).double().mean().item(), " (u64 > 1e-30 only):",
              ((allrel > 1e-4) & (u64 > 1e-30)).double().mean().item())
ref_a = (-(ux64 + 1e-35).log()).sum(1).cpu()
print("out_a fp64 ", ref_a.tolist())
print("out_a simt ", outs[False][0].tolist())
print("out_a tc   ", outs[True][0].tolist())
print("--- per kind: max relative difference tc vs CUDA-core per output (a, b, c, d, nll) and of the logit gradient")
for kind, kw in ((nat.LOSS_SDDM, dict(xt=xtil)), (nat.LOSS_SDDM, dict(xt=xtil, logit_branch=nat.BRANCH_SDDM_REVERSE_LOGSCALE)),
                 (nat.LOSS_CRM, dict(xt=xt, crm_type=0)), (nat.LOSS_CRM, dict(xt=xt, crm_type=1)), (nat.LOSS_CRM, dict(xt=xt, crm_type=2))):
    res = []
    for tc in (False, True):
        ops._LossTerms.use_tc = tc
        lg = logits.clone().requires_grad_(True)
        o = ops.loss_terms(lg, kind, Q=Q, QT=QT, Rb=Rb, beta=beta, x0=x0, eps=1e-9, **kw)
        (o[0].sum() + 0.3 * o[1].sum() + 0.1 * o[3].sum() + 0.01 * o[4].sum()).backward()
        res.append(([t.detach().cpu().double() for t in o], lg.grad.detach().cpu().double()))
    ops._LossTerms.use_tc = True
    rels = [((a - b).abs() / a.abs().clamp_min(1e-6)).max().item() for a, b in zip(res[0][0], res[1][0])]
    gmax = res[0][1].abs().max().item()
    print(kind, {k: v for k, v in kw.items() if k != "xt"}, ["%.2e" % r for r in rels], "grad %.2e of max" % ((res[0][1] - res[1][1]).abs().max().item() / gmax))
