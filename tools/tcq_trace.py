"""Diagnostic (not part of the product): per-tile timeline of the step_q_kernel warp roles (first CTA pair).
Build the traced variant first:  python tools/variants.py build trace ctdd_step_tcq.cu CTDD_TC_TRACE
and run with CTDD_B200_LIB=<package>/build/variants/libctdd_trace.so."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ctdd_b200 import _native as nat, make_config, ops
from ctdd_b200.lib.models import forward_model as fm

w = bench.WORKLOADS["C4"]
B = int(os.environ.get("B", 1024)); D, S = w["D"], w["S"]
dev = torch.device("cuda")
cfg = make_config(data=dict(S=S), model=dict(w["model"]), device="cuda")
model = fm.GaussianTargetRate(cfg, "cuda")
ts = [float(v) for v in os.environ.get("TS", "0.7,0.1").split(",")]
Q, QT, beta = model.qt0_tables(ts, dev)
Rb, RbT = model.base_rate_tables(dev)
branch = nat.BRANCH_TAULDR
tc = ops.prep_tc_tables(Q, QT, Rb, 1e-9, branch); tcs = ops.prep_tc_static(Rb)
lg, x0 = bench.synth_logits(B, D, S, 1, dev)
x = torch.clamp(x0 + torch.randint(-3, 4, x0.shape, device=dev), 0, S - 1).to(torch.int32)
L = nat.lib()
L.ctdd_debug_trace_read_q.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
sl = slice(40, 300)
med = lambda v: float(np.median(v))
for i, t in enumerate(ts):
    for rep in range(2):
        L.ctdd_debug_trace_clear_q()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.reverse_step(nat.MODE_TAU_LEAP, branch, lg, x, Q[i], QT[i], Rb, RbT, beta[i], 0.99 / 1000, 1e-9, N=B, D=D, S=S,
                         seed=1, offset=i, tc_tables=tc[i], tc_static=tcs)
        e1.record(); torch.cuda.synchronize()
    tr = np.zeros((2, 7, 512, 8), dtype=np.int64)
    assert L.ctdd_debug_trace_read_q(tr.ctypes.data, tr.nbytes) == 0
    np.save(f"gpurun_out/tcq_trace_t{t}.npy", tr)
    print(f"=== t={t} kernel {e0.elapsed_time(e1):.3f} ms; first stamp -> last stamp of CTA 0: {int(tr[0][tr[0] > 0].max() - tr[0][tr[0] > 0].min())} cycles")
    for cta in range(2):
        p, m, f = tr[cta, 0], tr[cta, 1], tr[cta, 4]
        print(f" CTA {cta} producer warp 0: wait empty {med(p[sl,1]-p[sl,0]):.0f}  first pass..arrive {med(p[sl,2]-p[sl,1]):.0f}  "
              f"lring waits/tile {med(p[sl,3]):.0f}  period {np.diff(p[sl,0]).mean():.0f}")
        if cta == 0:
            print(f"       MMA: wait full {med(m[sl,1]-m[sl,0]):.0f}  wait tmem_empty {med(m[sl,2]-m[sl,1]):.0f}  issue {med(m[sl,3]-m[sl,2]):.0f}  period {np.diff(m[sl,0]).mean():.0f}")
        for role in (2, 3):
            e = tr[cta, role]
            print(f"       epi h={role-2}: wait scal {med(e[sl,1]-e[sl,0]):.0f}  tmem_full {med(e[sl,2]-e[sl,1]):.0f}  cfree {med(e[sl,3]-e[sl,2]):.0f}  "
                  f"ld0 {med(e[sl,4]-e[sl,3]):.0f}  work0 {med(e[sl,5]-e[sl,4]):.0f}  ld1 {med(e[sl,6]-e[sl,5]):.0f}  work1 {med(e[sl,7]-e[sl,6]):.0f}  "
                  f"mma commit->tmem_full seen {med(e[sl,2]-tr[0,1][sl,3]):.0f}  period {np.diff(e[sl,0]).mean():.0f}")
        d, pp = tr[cta, 5], tr[cta, 6]
        print(f"       epi detail (h=0, batch 0): after ld_wait->transposed {med(d[sl,1]-tr[cta,2][sl,4]):.0f}  ->table row {med(d[sl,2]-d[sl,1]):.0f}  "
              f"->prefix chain {med(d[sl,3]-d[sl,2]):.0f}  ->counts {med(d[sl,4]-d[sl,3]):.0f}  ->picks {med(d[sl,5]-d[sl,4]):.0f}  rows with K>0: {d[sl,6].mean():.2f}")
        print(f"       producer pass detail: ring wait {med(pp[sl,1]-pp[sl,0]):.0f}  ->values {med(pp[sl,2]-pp[sl,1]):.0f}  ->refill issued {med(pp[sl,3]-pp[sl,2]):.0f}  "
              f"->row max {med(pp[sl,4]-pp[sl,3]):.0f}  ->finish_prev {med(pp[sl,5]-pp[sl,4]):.0f}  ->(empty wait) exp+sums {med(pp[sl,6]-pp[sl,5]):.0f}  "
              f"->split+stores issued {med(pp[sl,7]-pp[sl,6]):.0f}  whole pass {med(pp[sl,7]-pp[sl,0]):.0f}")
        print(f"       finalizer: wait scal {med(f[sl,1]-f[sl,0]):.0f}  wait contrib {med(f[sl,2]-f[sl,1]):.0f}  work {med(f[sl,3]-f[sl,2]):.0f}")
