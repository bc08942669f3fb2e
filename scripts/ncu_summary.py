#!/usr/bin/env python
"""Condense an .ncu-rep (read here, no GPU needed) into the per-kernel numbers the roofline is argued from.

    python scripts/ncu_summary.py gpurun_out/prof_scan.ncu-rep > profiles/r01_scan_full.txt
    python scripts/ncu_summary.py --launches gpurun_out/launches.csv > profiles/r01_launches.txt
"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
]


def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ki = hdr.index("Kernel Name")
    for row in rows[2:]:
        print(f"== {row[ki][:110]}")
        for w, i in idx:
            print(f"   {w:72s} {row[i]} {units[i]}")
        try:
            rd = float(row[hdr.index('dram__bytes_read.sum')]); wr = float(row[hdr.index('dram__bytes_write.sum')])
            print(f"   {'traffic = dram read + write':72s} {rd:.3f} {units[hdr.index('dram__bytes_read.sum')]} + {wr:.3f} {units[hdr.index('dram__bytes_write.sum')]}")
        except Exception:
            pass


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(io.StringIO("".join(lines)))
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for row in r:
        v = float(row[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row[ui], v)
        a = agg.setdefault(row[ki][:100], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print(f"# per-launch device time (gpu__time_duration.sum, --clock-control none): cold-cache, serialised -> compare SHARES")
    print(f"# {'total us':>12s} {'share':>7s} {'launches':>8s} {'us/launch':>10s}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {t:12.1f} {100 * t / tot:6.1f}% {c:8d} {t / c:10.1f}  {n}")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[1])
