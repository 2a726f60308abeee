"""Diagnostic (not part of the product): join an ncu SASS source page with nvdisasm line info and aggregate the
executed warp-instructions / stall samples of the tcgen05 kernel per source line and per warp role.

  ncu -i rep.ncu-rep --page source --csv > src.csv
  python tools/ncu_lines.py src.csv '<mangled kernel name>' [lo-hi:label ...]
"""
import csv, re, subprocess, sys, os, collections, tempfile

SRC = os.environ.get("CTDD_SRC", "ctdd_step_tcq")   # source file (without .cu) whose kernel is analysed
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "continuous-time-diffusion-models-for-discrete-data_b200", "libctdd_b200.so")


def line_map(kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", SRC + ".sm_100a.cubin", LIB], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    m, cur, infn = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("\t.text.") or ln.startswith(".text."):
            infn = kernel in ln
        if not infn:
            continue
        f = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if f:
            # keep the OUTERMOST location in ctdd_step_tc.cu (inlined helpers report their own line first)
            if f.group(1).endswith(SRC + ".cu") and "inlined at" not in f.group(3):
                cur = int(f.group(2))
            elif "inlined at" in f.group(3):
                g = re.findall(r'inlined at "([^"]+)", line (\d+)', f.group(3))
                g = [int(b) for a, b in g if a.endswith(SRC + ".cu")]
                if g:
                    cur = g[-1]
            continue
        a = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if a and cur is not None:
            m[int(a.group(1), 16)] = (cur, a.group(2).strip())
    return m


def main():
    src, kernel = sys.argv[1], sys.argv[2]
    ranges = []
    for a in sys.argv[3:]:
        r, label = a.split(":")
        lo, hi = r.split("-")
        ranges.append((int(lo), int(hi), label))
    lm = line_map(kernel)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    iA, iI, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    per_line = collections.defaultdict(lambda: [0, 0])
    per_role = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = 0
    for r in rows[2:]:
        try:
            addr = int(r[iA], 16)
        except Exception:
            continue
        if base is None:
            base = addr
        off = addr - base
        n, s = int(r[iI]), int(r[iS])
        line = lm.get(off, (0, "?"))[0]
        per_line[line][0] += n
        per_line[line][1] += s
        tot += n
        label = next((lb for lo, hi, lb in ranges if lo <= line <= hi), "other")
        per_role[label][0] += n
        per_role[label][1] += s
        for i in stall_cols:
            v = int(r[i] or 0)
            if v:
                per_role[label][2][hdr[i]] += v
    print("total warp instructions", tot)
    for lb, (n, s, st) in sorted(per_role.items(), key=lambda kv: -kv[1][0]):
        top = ", ".join(f"{k[6:]} {v * 100 // max(s, 1)}%" for k, v in st.most_common(5))
        print(f"  {lb:12s} inst {n:12d} ({100 * n / tot:5.1f}%)  samples {s:8d}  [{top}]")
    print("top lines:")
    for line, (n, s) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"  line {line:4d}  inst {n:11d} ({100 * n / tot:4.1f}%)  samples {s}")


if __name__ == "__main__":
    main()
