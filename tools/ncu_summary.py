"""Summarise an ncu launch list (csv of gpu__time_duration.sum) by kernel family and by (kernel, grid)."""
import collections
import csv
import re
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("<unnamed>::", "")
    return n[:64]


fam, agg = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    us = float(r["Metric Value"]) / 1e3
    k = short(r["Kernel Name"])
    fam[k.split("<")[0]][0] += 1
    fam[k.split("<")[0]][1] += us
    agg[k + " grid=" + r["Grid Size"]][0] += 1
    agg[k + " grid=" + r["Grid Size"]][1] += us
tot = sum(v[1] for v in fam.values())
print(f"{len(rows)} launches, {tot:.1f} us of kernel time (serialised, per-launch cold-cache durations)")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[0]:5d} {v[1]:9.1f} us {100 * v[1] / tot:5.1f} % {v[1] / v[0]:7.1f} avg  {k}")
print()
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[0]:5d} {v[1]:9.1f} us {v[1] / v[0]:7.1f} avg  {k}")
