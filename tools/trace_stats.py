"""Diagnostic: median per-segment cycle counts from a tc_trace.npy dump (tools/tc_trace.py)."""
import sys
import numpy as np
tr = np.load(sys.argv[1])
sl = slice(100, 600)
for cta in range(2):
    p = tr[cta, 0]
    names = ["top->empty_ok", "empty_ok->ring0", "ring0->pass0end", "pass0end->ring1", "ring1->reduce1", "reduce1->pass1end",
             "pass1end->arrive", "arrive->next top"]
    segs = [p[sl, 1] - p[sl, 0], p[sl, 5] - p[sl, 1], p[sl, 2] - p[sl, 5], p[sl, 6] - p[sl, 2], p[sl, 7] - p[sl, 6],
            p[sl, 3] - p[sl, 7], p[sl, 4] - p[sl, 3], p[101:601, 0] - p[sl, 4]]
    print("CTA", cta, "producer:", " ".join(f"{n}={np.median(x):.0f}" for n, x in zip(names, segs)), "period", round(np.diff(p[sl, 0]).mean()))
    for role, nm in ((2, "epi_L"), (3, "epi_R")):
        e = tr[cta, role]
        n2 = ["A:wait tmem_full", "A:wait gfree", "A:ld+store", "A:arrive", "S:wait", "S:sample"]
        s2 = [e[sl, 1] - e[sl, 0], e[sl, 2] - e[sl, 1], e[sl, 3] - e[sl, 2], e[sl, 4] - e[sl, 3], e[sl, 6] - e[sl, 5], e[sl, 7] - e[sl, 6]]
        print("   ", nm, " ".join(f"{n}={np.median(x):.0f}" for n, x in zip(n2, s2)))
m = tr[0, 1]
print("MMA: wait full", np.median(m[sl, 1] - m[sl, 0]), "wait tmem_empty", np.median(m[sl, 2] - m[sl, 1]), "issue",
      np.median(m[sl, 3] - m[sl, 2]), "period", np.diff(m[sl, 0]).mean())
