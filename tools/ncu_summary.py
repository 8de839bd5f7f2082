#!/usr/bin/env python
"""Condense ncu output into the small text summaries committed under profiles/.

  tools/ncu_summary.py launches <launches.csv> [steps]   -> per-kernel share of device time (gpu__time_duration.sum)
  tools/ncu_summary.py full <report.ncu-rep>             -> per-launch key metrics of a `--set full` capture
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
        agg[n][0] += 1
        agg[n][1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none ; %d launches, %.2f ms total device time" %
          (len(rows), tot))
    print("# (cold-cache, serialised: compare SHARES with bench.py's live CUDA-event numbers, not absolutes)")
    print("%10s %7s %7s  %s" % ("ms", "share", "count", "kernel"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.3f %6.1f%% %7d  %s" % (v[1], 100 * v[1] / tot, v[0], k[:120]))


FULL_METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_smem"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(m), lab, units[hdr.index(m)]) for m, lab in FULL_METRICS if m in hdr]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    print("# ncu --set full --clock-control none : %s" % path)
    print("  ".join(["%-34s %-16s" % ("kernel", "grid")] + ["%12s" % ("%s[%s]" % (l, u)) for _, l, u in cols]))
    for d in data:
        print("  ".join(["%-34s %-16s" % (re.sub(r"\(.*", "", d[ki])[:34], d[gi])] + ["%12s" % d[i][:12] for i, _, _ in cols]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2])
