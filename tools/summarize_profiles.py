#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep + launches.csv into the small tracked summaries under profiles/ (run here, no GPU)."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    os.makedirs(OUT, exist_ok=True)
    res = {}
    for name in ("prof_k1", "prof_k1_emit", "prof_gram"):
        rep = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        hdr, units, data = raw(rep)
        with open(os.path.join(OUT, "%s_ncu_%s.csv" % (TAG, name[5:])), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(data))])
            w.writerow(["Kernel Name", ""] + [r[hdr.index("Kernel Name")][:60] for r in data])
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    w.writerow([k, units[i]] + [r[i] for r in data])
        i_r, i_w, i_t = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")

        def tobytes(v, u):
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        res[name] = [{"dram_bytes": tobytes(r[i_r], units[i_r]) + tobytes(r[i_w], units[i_w]), "time": r[i_t] + " " + units[i_t]} for r in data]
    if "prof_k1" in res:
        json.dump({"kernel": "k1_dense_kernel steady-state pass (no emit), BASELINE configs[1] at N=1 (8 partitions resident)",
                   "traffic_bytes_per_launch": res["prof_k1"][0]["dram_bytes"], "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum",
                   "launches": res["prof_k1"]}, open(os.path.join(OUT, "%s_k1_traffic.json" % TAG), "w"), indent=1)
    lc = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        t, c = defaultdict(float), defaultdict(int)
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            nm = r[ki].split("(")[0].replace("void ", "")
            t[nm] += v
            c[nm] += 1
        mine = {k: v for k, v in t.items() if k.startswith("mlease::") or k.startswith("<unnamed>::")}
        tot = sum(mine.values())
        with open(os.path.join(OUT, "%s_launch_shares.csv" % TAG), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_ns", "share_of_library_kernels"])
            for k, v in sorted(mine.items(), key=lambda x: -x[1]):
                w.writerow([k, c[k], int(v), "%.4f" % (v / tot)])
            w.writerow(["# command: ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --no-e2e --no-cpu --steps 3 --warmup 1 (torch data-generation kernels excluded)", "", "", ""])
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
