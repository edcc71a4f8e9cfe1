"""Summarise an ncu launch list (csv of gpu__time_duration.sum [+ dram__bytes_read/write.sum]) by kernel family and by
(kernel, grid).  `--json out.json` also writes the per-family averages bench.py reads for `roofline.traffic`."""
import collections
import csv
import json
import re
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rows = list(csv.DictReader(l for l in open(args[0]) if not l.startswith("==")))
topn = int(args[1]) if len(args) > 1 else 30


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("<unnamed>::", "")
    return n[:64]


def unit_scale(u):
    return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


launch = collections.OrderedDict()  # ID -> {name, grid, us, rd, wr}
for r in rows:
    d = launch.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "grid": r["Grid Size"], "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", "")) * unit_scale(r["Metric Unit"])
    if r["Metric Name"].startswith("gpu__time"):
        d["us"] = v
    elif "read" in r["Metric Name"]:
        d["rd"] = v
    elif "write" in r["Metric Name"]:
        d["wr"] = v
fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in launch.values():
    for table, key in ((fam, d["name"].split("<")[0]), (agg, d["name"] + " grid=" + d["grid"])):
        table[key][0] += 1
        table[key][1] += d["us"]
        table[key][2] += d["rd"] + d["wr"]
tot = sum(v[1] for v in fam.values())
have_dram = any(v[2] for v in fam.values())
print(f"{len(launch)} launches, {tot:.1f} us of kernel time (serialised by ncu, per-launch cold-cache durations)")
print("count   total_us   share   avg_us" + ("   avg_dram_MB  dram_GB/s" if have_dram else "") + "  kernel")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    extra = f" {v[2] / v[0] / 1e6:12.2f} {v[2] / v[1] / 1e3:10.0f}" if have_dram else ""
    print(f"{v[0]:5d} {v[1]:10.1f} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:8.1f}{extra}  {k}")
print()
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    extra = f" {v[2] / v[0] / 1e6:12.2f} {v[2] / v[1] / 1e3:10.0f}" if have_dram else ""
    print(f"{v[0]:5d} {v[1]:10.1f} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:8.1f}{extra}  {k}")
if "--json" in sys.argv:
    out = {k: {"launches": v[0], "us_total": round(v[1], 1), "dram_bytes_per_launch": round(v[2] / v[0]) if have_dram else None}
           for k, v in fam.items()}
    json.dump({"source": args[0], "serialised_us": round(tot, 1), "families": out}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
