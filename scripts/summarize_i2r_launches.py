"""Per-kernel summary of an ncu launch list of `bench.py --workload img2refmap` (last launch sequence in the file).
usage: summarize_i2r_launches.py <csv> [out.txt]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, x in enumerate(rows) if x and x[0] == "ID"][0]
H = rows[hdr]
ki, vi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
cur = {}
for x in rows[hdr + 1:]:
    if len(x) > vi:
        cur.setdefault((int(x[0]), x[ki].split("(")[0]), {})[x[mi]] = float(x[vi].replace(",", ""))
ids = sorted(cur)
last_hist = max(k[0] for k in ids if "hist_pass" in k[1])
lines, tot, dram = [], 0.0, 0.0
for k in ids:
    if k[0] < last_hist:
        continue
    m = cur[k]
    t = m.get("gpu__time_duration.sum", 0) / 1e3
    d = (m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0))
    d = d / 1e6 if d > 1e4 else d  # ncu prints MB or bytes depending on the magnitude
    tot += t; dram += d
    lines.append(f"drm::{k[1]:28s} {t:8.1f} us  dram {d:8.1f} MB  {m.get('smsp__inst_executed.sum', 0) / 1e6:7.1f} M warp-instr")
lines.append(f"total {tot:.1f} us, DRAM {dram:.1f} MB")
print("\n".join(lines))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(lines) + "\n")
