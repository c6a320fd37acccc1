"""Per-kernel summary of an ncu launch list (the CSV written by
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
 --profile-from-start off --csv --log-file <csv> python bench.py --ncu`):
launches, total / average duration, share of the serialised step, DRAM bytes and GB/s per kernel.
usage: python scripts/summarize_launches.py <launches.csv> [title]"""
import csv, re, sys
from collections import OrderedDict

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ix = {n: i for i, n in enumerate(hdr)}
per = OrderedDict()
for r in rows[1:]:
    key = r[ix["ID"]]
    d = per.setdefault(key, {"name": r[ix["Kernel Name"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("u") else v * 1e3)
    elif m.startswith("dram__bytes"):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d["bytes"] = d.get("bytes", 0.0) + v * scale
agg = {}
for d in per.values():
    name = re.sub(r"\(.*", "", d["name"]).strip()
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("us", 0.0)
    a[2] += d.get("bytes", 0.0)
total = sum(a[1] for a in agg.values())
print(f"{title}; total {total / 1e3:.1f} ms serialized over {sum(a[0] for a in agg.values())} launches")
for name, (n, us, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:58]:58s} n={n:4d} total={us / 1e3:8.2f} ms share={100 * us / total:5.1f}% avg={us / n:9.1f} us  "
          f"dram={b / 1e9:7.2f} GB ({b / max(us, 1e-9) / 1e6:5.2f} TB/s)")
