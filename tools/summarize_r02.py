#!/usr/bin/env python
"""Round-2 profile summaries: turns scratch files under gpurun_out/ into the small tracked files under profiles/ (run here, no GPU).

  python tools/summarize_r02.py ncu   <rep under gpurun_out> <out csv name>       ncu --set full capture -> the metrics that matter
  python tools/summarize_r02.py list  <launch csv under gpurun_out> <out csv name> launch list -> per-kernel counts / time / share
  python tools/summarize_r02.py traffic <rep> <workload cfgN> <kernel label> [algorithmic bytes of the captured launch]   -> profiles/k1_traffic.json
  python tools/summarize_r02.py sass                                                per-kernel tcgen05 / TMA / DMMA mnemonic counts
"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[1], rows[2:]


def tobytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def cmd_ncu(rep, out):
    hdr, units, data = raw(os.path.join(GO, rep))
    with open(os.path.join(OUT, out), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(data))])
        w.writerow(["Kernel Name", ""] + [r[hdr.index("Kernel Name")][:70] for r in data])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in data])
    print("wrote", out)


def cmd_traffic(rep, workload, label, alg_bytes=None):
    hdr, units, data = raw(os.path.join(GO, rep))
    i_r, i_w, i_t = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    path = os.path.join(OUT, "k1_traffic.json")
    j = json.load(open(path)) if os.path.exists(path) else {}
    launches = [{"dram_bytes": tobytes(r[i_r], units[i_r]) + tobytes(r[i_w], units[i_w]), "time": r[i_t] + " " + units[i_t]} for r in data]
    j[workload] = {"kernel": label, "traffic_bytes_per_launch": launches[0]["dram_bytes"],
                   "algorithmic_bytes_of_captured_launch": float(alg_bytes) if alg_bytes else None,
                   "traffic_over_algorithmic": (launches[0]["dram_bytes"] / float(alg_bytes)) if alg_bytes else None,
                   "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, capture gpurun_out/%s summarised in profiles/" % rep,
                   "launches": launches}
    json.dump(j, open(path, "w"), indent=1)
    print("wrote k1_traffic.json", workload, launches[0])


def cmd_list(src, out):
    lines = [ln for ln in open(os.path.join(GO, src)) if not ln.startswith("==")]
    t, c = defaultdict(float), defaultdict(int)
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1)
        nm = row["Kernel Name"].split("(")[0].replace("void ", "")
        t[nm] += v
        c[nm] += 1
    tot = sum(t.values())
    with open(os.path.join(OUT, out), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "avg_us", "share_of_library_kernels"])
        for k, v in sorted(t.items(), key=lambda x: -x[1]):
            w.writerow([k, c[k], "%.3f" % (v / 1e6), "%.1f" % (v / c[k] / 1e3), "%.4f" % (v / tot)])
        w.writerow(["# ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mlease (serialised, cold cache: shares, not absolutes)", "", "", "", ""])
    print("wrote", out, "total ms %.1f" % (tot / 1e6))


def cmd_sass():
    so = os.path.join(ROOT, "ml-ease_b200", "lib", "libmlease_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    pats = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "DMMA", "HMMA", "ATOMS", "SYNCS", "UTCATOMSWS"]
    cur, cnt = None, defaultdict(lambda: defaultdict(int))
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        if cur:
            for p in pats:
                if re.search(r"\b" + p, ln):
                    cnt[cur][p] += 1
    with open(os.path.join(OUT, "r02_sass_summary.txt"), "w") as f:
        f.write("cuobjdump -sass ml-ease_b200/lib/libmlease_b200.so : occurrences of the Blackwell mnemonics per kernel\n")
        f.write("(UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4, UTMALDG = TMA tensor load, UBLKCP = bulk TMA copy, LDTM = tcgen05.ld,\n")
        f.write(" DMMA = fp64 mma.sync, HMMA = mma.sync m16n8k8 tf32 (the TF32 merges of the wide inverse), ATOMS = shared-memory atomics, SYNCS = mbarrier ops)\n\n")
        for k in sorted(cnt):
            if cnt[k]:
                f.write("%-75s %s\n" % (k[:75], "  ".join("%s x%d" % (p, n) for p, n in sorted(cnt[k].items()))))
    print("wrote r02_sass_summary.txt")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    {"ncu": cmd_ncu, "list": cmd_list, "traffic": cmd_traffic, "sass": lambda: cmd_sass()}[sys.argv[1]](*sys.argv[2:])
