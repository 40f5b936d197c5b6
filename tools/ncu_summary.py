#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list CSV (gpu__time_duration) and/or a
--set full .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def launches(path):
  rows = [r for r in csv.DictReader(l for l in open(path) if not l.startswith("=="))]
  agg = collections.OrderedDict()
  for r in rows:
    name = r["Kernel Name"].split("(")[0][:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", ""))
  tot = sum(v[1] for v in agg.values())
  print("## launch list: %s (%d launches, %.3f ms total device time, cold-cache serialised)" % (path, len(rows), tot / 1e6))
  print("| kernel | launches | total ms | avg us | share |")
  print("|---|---|---|---|---|")
  for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f | %.1f%% |" % (n, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / tot))


def full(path):
  out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  print("## full capture: %s" % path)
  for r in rows[2:]:
    print("### `%s` grid %s block %s" % (r[hdr.index("Kernel Name")][:80], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
    print("| metric | value | unit |")
    print("|---|---|---|")
    for k in KEYS:
      if k in hdr:
        print("| %s | %s | %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    for i, k in enumerate(hdr):
      if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
        try:
          v = float(r[i])
        except ValueError:
          continue
        if v >= 0.05:
          print("| stall:%s | %.3f | warps/issue |" % (k[len(STALL):-len("_per_issue_active.ratio")], v))


if __name__ == "__main__":
  for p in sys.argv[1:]:
    (full if p.endswith(".ncu-rep") else launches)(p)
    print()
