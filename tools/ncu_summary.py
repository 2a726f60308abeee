"""Read an `ncu --set full` report (first profiled kernel) and (a) print its key metrics as a markdown table row set,
(b) record the DRAM traffic per launch in profiles/ncu_traffic.json under a workload key - bench.py reports that number as
roofline.traffic.   python tools/ncu_summary.py gpurun_out/x.ncu-rep C4 ["note"]      (runs here, no GPU needed)"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, workload = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    got = {}
    for k in KEYS:
        for i, h in enumerate(hdr):
            if h == k:
                got[k] = (vals[i], units[i])
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"kernel: {kname}")
    for k, (v, u) in got.items():
        print(f"| {k} | {v} {u} |")
    rd = float(got["dram__bytes_read.sum"][0].replace(",", "")) * UNIT.get(got["dram__bytes_read.sum"][1], 1.0)
    wr = float(got["dram__bytes_write.sum"][0].replace(",", "")) * UNIT.get(got["dram__bytes_write.sum"][1], 1.0)
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    d[workload] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": kname, "report": os.path.basename(rep),
                   "duration_ms_under_ncu": float(got["gpu__time_duration.sum"][0].replace(",", "")) * (1e-3 if got["gpu__time_duration.sum"][1] == "us" else 1.0),
                   "note": note}
    json.dump(d, open(p, "w"), indent=1, sort_keys=True)
    print("traffic", rd + wr, "->", p)


if __name__ == "__main__":
    main()
