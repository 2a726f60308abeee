"""Every kernel family of libctdd_b200.so once at its configuration size: CUDA-event time, algorithmic
bytes / FLOP and the resulting GB/s / TFLOP/s, one JSON line per kernel.  Not part of the product.

    python tools/kernel_zoo.py                      # timings (3 warm-ups, mean of 10)
    ZOO_NCU=1 ncu --set full --clock-control none -k regex:'<names>' python tools/kernel_zoo.py    # one launch each
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ctdd_b200 import _native as nat, make_config, ops  # noqa: E402
from ctdd_b200.lib.models import forward_model as fm  # noqa: E402

NCU = bool(os.environ.get("ZOO_NCU"))
dev = torch.device("cuda")
PEAK_HBM, PEAK_FMA = 6529.0, 148 * 128 * 2 * 1.965e-3      # GB/s measured (MEASURED_PEAKS.json), TFLOP/s fp32 FMA nominal
try:
    PEAK_HBM = float(json.load(open(os.path.join(os.path.dirname(bench.__file__), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=10, warm=3):
    if NCU:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, what, ms, bytes_=None, flop=None, note=""):
    rec = {"kernel": name, "workload": what, "ms": round(ms, 5)}
    if bytes_ is not None:
        rec["alg_bytes"] = int(bytes_)
        rec["GB/s"] = round(bytes_ / ms / 1e6, 1)
        rec["frac_hbm_peak"] = round(bytes_ / ms / 1e6 / PEAK_HBM, 3)
    if flop is not None:
        rec["alg_flop"] = int(flop)
        rec["TFLOP/s"] = round(flop / ms / 1e9, 2)
    if note:
        rec["note"] = note
    print(json.dumps(rec), flush=True)


def model_for(wname):
    w = bench.WORKLOADS[wname]
    cfg = make_config(data=dict(S=w["S"]), model=dict(w["model"], Q_sigma=w["model"].get("Q_sigma", 20.0)), device="cuda")
    return w, getattr(fm, bench.MIXIN[w["fwd"]])(cfg, "cuda")


def steps():
    # C1 / C2 also at 16 x the configuration's batch (the HBM-rate claim of the small-S kernels is made there;
    # at the configurations' own batch the inputs fit in L2 and back-to-back Python calls time the host)
    for wname, t, mult in (("C1", 0.5, 1), ("C1", 0.5, 16), ("C2", 0.1, 1), ("C2", 0.1, 16), ("C3", 0.5, 1)):
        w, model = model_for(wname)
        S, D, B = w["S"], w["D"], w["B"] * mult
        Q, QT, beta = model.qt0_tables([t], dev)
        Rb, RbT = model.base_rate_tables(dev)
        branch = nat.branch_for(w["loss"], None)
        mode = nat.MODE_EULER if w["mode"] == "euler" else nat.MODE_TAU_LEAP
        tc = ops.prep_tc_tables(Q, QT, Rb, 1e-9, branch) if S == 256 else None
        tcs = ops.prep_tc_static(Rb) if tc is not None else None
        lg, x0 = bench.synth_logits(B, D, S, 7, dev, None)
        x = x0.to(torch.int32)
        ws = torch.empty((max(1, int(nat.lib().ctdd_step_workspace_bytes(B * D, S, 0))),), dtype=torch.uint8, device=dev)
        h = (w["max_t"] - w["min_t"]) / w["num_steps"]

        def run():
            ops.reverse_step(mode, branch, lg, x, Q[0], QT[0], Rb, RbT, beta[0], h, 1e-9, N=B, D=D, S=S,
                             reject_multi=not w["ordinal"], seed=1, offset=0, tc_tables=(tc[0] if tc is not None else None),
                             tc_static=tcs, workspace=ws)
        ms = timed(run)
        name = "step_tc_kernel" if S == 256 else f"step_small_kernel<{S}>"
        report(name, f"{wname}: S={S} D={D} N={B} {w['mode']} t={t}", ms, bytes_=4.0 * B * D * S + 8.0 * B * D,
               flop=2.0 * B * D * S * S if S == 256 else None)
    # CUDA-core block kernel at S = 256 (SDDM direct branch / cross-check path), and the exact-posterior branch sizes
    w, model = model_for("C3")
    S, D, B = 256, 784, 64
    Q, QT, beta = model.qt0_tables([0.5], dev)
    Rb, RbT = model.base_rate_tables(dev)
    lg, x0 = bench.synth_logits(B, D, S, 8, dev, None)
    x = x0.to(torch.int32)

    def run_block():
        ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, lg, x, Q[0], QT[0], Rb, RbT, beta[0], 1e-3, 1e-9, N=B, D=D, S=S,
                         seed=1, offset=0, impl=nat.IMPL_SIMT)
    report("step_block_kernel", f"S=256 D=784 N={B} tau_leap (CUDA-core path)", timed(run_block), bytes_=4.0 * B * D * S + 8.0 * B * D,
           flop=2.0 * B * D * S * S)


def forward_process():
    w, model = model_for("C4")
    S, D = 256, 3072
    for B in (128, 1024):
        ts = torch.rand(B, device=dev) * 0.98 + 0.01

        def run_q():
            model._build_qt0(model._transition_delta(ts), inverse=True, want_transpose=True)
        report("qt0_fused_kernel", f"q_t|0 for B={B} distinct times, S=256 (Q and Q^T)", timed(run_q),
               bytes_=8.0 * B * S * S, flop=2.0 * B * S * S * S)
    B = 128
    ts = torch.rand(B, device=dev) * 0.98 + 0.01
    Q, QT = model._build_qt0(model._transition_delta(ts), inverse=True, want_transpose=True)
    beta = model._rate_scalar(ts).float().contiguous()
    Rb, _ = model.base_rate_tables(dev)
    x0 = torch.randint(0, S, (B, D), device=dev, dtype=torch.int32)
    report("noise_xt_kernel + xtilde_kernel", f"x_t ~ q_t|0(.|x0) and x~, B={B} D={D} S=256", timed(lambda: ops.noise_xt(Q, Rb, beta, x0, 1, 0)),
           bytes_=12.0 * B * D, note="row gathers of Q[b, x0, :] come from L2 (B * 256 KB of Q)")
    probs = torch.softmax(-((torch.arange(S, device=dev) - 128.0) / 40.0) ** 2, 0).contiguous()
    rows = 1024 * D
    report("categorical_shared_kernel", f"initial states, {rows} rows, S=256", timed(lambda: ops.sample_categorical_shared(probs, rows, 1)),
           bytes_=4.0 * rows)


def losses():
    w, model = model_for("C4")
    S = 256
    Rb, _ = model.base_rate_tables(dev)
    for name, kind, B, D in (("C3 CatRMNLL", nat.LOSS_CRM, 64, 784), ("C5 SDDMElbo", nat.LOSS_SDDM, 64, 3072),
                             ("C5 CTElbo", nat.LOSS_CTELBO, 64, 3072)):
        ts = torch.rand(B, device=dev) * 0.98 + 0.01
        Q, QT = model._build_qt0(model._transition_delta(ts), inverse=True, want_transpose=True)
        beta = model._rate_scalar(ts).float().contiguous()
        x0 = torch.randint(0, S, (B, D), device=dev, dtype=torch.int32)
        xt, xtil = ops.noise_xt(Q, Rb, beta, x0, 1, 0)
        logits = (torch.randn(B, D, S, device=dev) - (torch.arange(S, device=dev).view(1, 1, S) - x0.unsqueeze(-1)) ** 2 / 128.0)
        logits.requires_grad_(True)

        def fwd():
            return ops.loss_terms(logits, kind, Q=Q, QT=QT, Rb=Rb, beta=beta, x0=x0, xt=xtil if kind != nat.LOSS_CRM else xt,
                                  x_tilde=xtil if kind == nat.LOSS_CTELBO else None, eps=1e-9)

        def fwd_bwd():
            sum(o.sum() for o in fwd()).backward()
            logits.grad = None
        with torch.no_grad():
            ms_f = timed(fwd)
        ms_fb = timed(fwd_bwd)
        report("loss_kernel<0>", f"{name} forward, B={B} D={D} S=256", ms_f, bytes_=4.0 * B * D * S, flop=2.0 * B * D * S * S)
        report("loss_kernel<0> + loss_kernel<1>", f"{name} forward + backward (torch autograd glue included)", ms_fb,
               bytes_=12.0 * B * D * S, flop=6.0 * B * D * S * S)


def head():
    B, D, S = 128, 3072, 256
    mu = torch.tanh(torch.randn(B, D, device=dev)).requires_grad_(True)
    ls = torch.randn(B, D, device=dev).requires_grad_(True)
    with torch.no_grad():
        report("logistic_logits_kernel", f"head forward B={B} D={D} S=256", timed(lambda: ops.logistic_logits(mu, ls, S, False)),
               bytes_=4.0 * B * D * S)
    g = torch.randn(B, D, S, device=dev)

    def fb():
        out = ops.logistic_logits_autograd(mu, ls, S, False)
        out.backward(g.view_as(out))
        mu.grad = ls.grad = None
    report("logistic_logits_kernel + logistic_backward_kernel", f"head forward + backward B={B} D={D} S=256", timed(fb),
           bytes_=8.0 * B * D * S)


if __name__ == "__main__":
    which = sys.argv[1:] or ["steps", "forward_process", "losses", "head"]
    for k in which:
        globals()[k]()
