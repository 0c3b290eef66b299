#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list per kernel name."""
import collections, csv, io, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.defaultdict(lambda: collections.defaultdict(float))
seen = collections.defaultdict(set)
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "")[:48]
    v = float(r["Metric Value"].replace(",", "")) * scale.get(r["Metric Unit"], 1.0)
    agg[name][r["Metric Name"]] += v
    seen[name].add(r["ID"])
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
print(f"{'kernel':48s} {'launches':>8s} {'time ms':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
for name, a in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
    print(f"{name:48s} {len(seen[name]):8d} {a['gpu__time_duration.sum']:10.4f} {100*a['gpu__time_duration.sum']/tot:6.1f}% "
          f"{a.get('dram__bytes_read.sum', 0):11.2f} {a.get('dram__bytes_write.sum', 0):11.2f}")
print(f"{'total':48s} {'':8s} {tot:10.4f}")
