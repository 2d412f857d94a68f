"""Summarise an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass) by hot regions."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "/tmp/prof_sass.csv"
per_step = float(sys.argv[2]) if len(sys.argv) > 2 else 4352 * 100  # warps * steps in the profiled launch
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 20
rows = list(csv.reader(open(path)))
hdr, data = rows[1], rows[2:]
ci, cs, csamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ci]) for r in data)
print("instructions per warp-step: %.1f   samples: %d" % (tot / per_step, sum(int(r[csamp]) for r in data)))


def level(n):
    x = n / per_step
    return 0 if x < 0.02 else 1 if x < 0.5 else 2 if x < 1.6 else 3


start, cur, regions = 0, level(int(data[0][ci])), []
for idx, r in enumerate(data[1:], 1):
    l = level(int(r[ci]))
    if l != cur:
        regions.append((start, idx - 1, cur))
        start, cur = idx, l
regions.append((start, len(data) - 1, cur))
for a, b, l in regions:
    s = sum(int(data[i][ci]) for i in range(a, b + 1))
    smp = sum(int(data[i][csamp]) for i in range(a, b + 1))
    if s / per_step > thr:
        ops = collections.Counter((data[i][cs].split()[1] if data[i][cs].strip().startswith('@') else data[i][cs].split()[0]).split('.')[0]
                                  for i in range(a, b + 1))
        print(f"rows {a}-{b} lvl{l} inst/warp-step {s / per_step:7.1f} samples {smp:6d}", ops.most_common(8))
