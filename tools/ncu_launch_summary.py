"""Per-kernel launch count / mean / total duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki].split("(")[0][:70]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):5d} mean={sum(v) / len(v):12.1f} {rows[1][ui]} share={sum(v) / tot:6.1%}")
