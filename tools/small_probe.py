"""Diagnostic: kernel time of the small-S reverse step (C1: S = 2 tau-leap, C2: S = 3 Euler) at the configuration's
batch and at 16 x that batch, as GB/s of algorithmic bytes (4*S + 8 per row).  Not part of the product."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ctdd_b200 import _native as nat, make_config, ops
from ctdd_b200.lib.models import forward_model as fm

dev = torch.device("cuda")
for wname, t in (("C1", 0.5), ("C2", 0.1)):
    w = bench.WORKLOADS[wname]
    S, D = w["S"], w["D"]
    cfg = make_config(data=dict(S=S), model=dict(w["model"], Q_sigma=20.0), device="cuda")
    model = getattr(fm, bench.MIXIN[w["fwd"]])(cfg, "cuda")
    Q, QT, beta = model.qt0_tables([t], dev)
    Rb, RbT = model.base_rate_tables(dev)
    branch = nat.branch_for(w["loss"], None)
    mode = nat.MODE_EULER if w["mode"] == "euler" else nat.MODE_TAU_LEAP
    if os.environ.get("MODE") == "corr":          # the corrector variants (general instantiation of the kernel)
        mode = nat.MODE_EULER_CORR if w["mode"] == "euler" else nat.MODE_TAU_LEAP_CORR
    h = (w["max_t"] - w["min_t"]) / w["num_steps"]
    for mult in (1, 16):
        B = w["B"] * mult
        lg, x0 = bench.synth_logits(B, D, S, 7, dev, None)
        x = x0.to(torch.int32)
        st = torch.zeros(8, dtype=torch.int64, device=dev)

        def run(stats=None):
            return ops.reverse_step(mode, branch, lg, x, Q[0], QT[0], Rb, RbT, beta[0], h, 1e-9, N=B, D=D, S=S,
                                    reject_multi=not w["ordinal"], seed=1, offset=0, stats=stats)
        out = run(st)
        with_stats = bool(os.environ.get("STATS"))       # STATS=1: the counters are accumulated in every timed call (the samplers' call)
        for _ in range(3):
            run(st if with_stats else None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            run(st if with_stats else None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        by = (4.0 * S + 8.0) * B * D
        s = st.cpu().numpy()
        print(f"{wname} S={S} rows={B*D:9d} ms={ms:8.4f} GB/s={by/ms/1e6:8.1f} rows/us={B*D/ms/1e3:8.1f} changed={s[0]/(B*D):.4f} "
              f"checksum={int(out['x'].long().sum())}")
