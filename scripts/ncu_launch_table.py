"""Print a per-launch table from an `ncu --csv` metrics log.  usage: python scripts/ncu_launch_table.py file.csv"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:44]), {})[r[mi]] = r[vi]
for (i, k), v in sorted(d.items()):
    t = float(v.get("gpu__time_duration.sum", "0").replace(",", "")) / 1e3
    rest = {a.split(".")[0].replace("smsp__", "").replace("sm__", ""): b for a, b in v.items() if a != "gpu__time_duration.sum"}
    print(f"{i:3d} {k:46s} {t:9.1f} us  {rest}")
