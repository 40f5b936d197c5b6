"""Per-source-line hot spots from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Line No")
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows:
  if len(r) > ii and r[0].isdigit() and r[2] == "-":
    try:
      data.append((int(r[si]), int(r[ii]), int(r[0]), r[1][:100]))
    except ValueError:
      pass
tot, toti = sum(d[0] for d in data), sum(d[1] for d in data)
print("total samples %d, warp instructions %d" % (tot, toti))
for d in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
  print("%5.1f%% smp %5.1f%% inst  L%d  %s" % (100 * d[0] / tot, 100 * d[1] / toti, d[2], d[3]))
