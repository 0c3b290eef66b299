"""Summarise an ncu launch list of `bench.py --workload render` (gpu__time_duration.sum [+ smsp__inst_executed.sum]) by
kernel and by footprint group of the last device-timed step.  usage: summarize_bench_launches.py <csv> [out.txt]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, x in enumerate(rows) if x and x[0] == "ID"][0]
H = rows[hdr]
ki, vi, mi, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name"), H.index("Grid Size")
cur = {}
for x in rows[hdr + 1:]:
    if len(x) <= vi:
        continue
    cur.setdefault((int(x[0]), x[ki].split("(")[0].replace("void ", ""), x[gi]), {})[x[mi]] = float(x[vi].replace(",", ""))
ids = sorted(cur)
tabs = [k[0] for k in ids if "tree_tables" in k[1]]
ngroups = 1
# one render call (one launch sequence) per step since the footprints share their launches; take the 2nd step (timed)
start, end = tabs[ngroups], tabs[2 * ngroups] if len(tabs) > 2 * ngroups else 10 ** 9
lines, tot, by = [], 0.0, {}
group = -1
for k in ids:
    if not (start <= k[0] < end) or not any(t in k[1] for t in ("drm::", "tree_", "pyr_", "render_")):
        continue
    m = cur[k]
    t = m.get("gpu__time_duration.sum", 0) / 1e6
    tot += t
    if "tree_tables" in k[1]:
        group += 1
        lines.append(f"-- render call {group} --")
    by[k[1]] = by.get(k[1], 0) + t
    if t > 0.2:
        lines.append(f"{k[1]:42s} grid {k[2]:18s} {t:9.3f} ms  {m.get('smsp__inst_executed.sum', 0) / 1e9:8.2f} G warp-instr")
lines.append(f"total of the step: {tot:.2f} ms (ncu-serialised launches; compare shares, not absolutes)")
for n, t in sorted(by.items(), key=lambda kv: -kv[1]):
    lines.append(f"  {n:42s} {t:9.3f} ms  {100 * t / tot:5.1f} %")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
