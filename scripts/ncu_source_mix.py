"""Executed-instruction mix of a kernel from an .ncu-rep (source page): per opcode the executed warp instructions and stall samples.
usage: python scripts/ncu_source_mix.py x.ncu-rep [kernel-regex] [top]"""
import csv, re, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, hdr, done = None, None, set()
ex, st = defaultdict(int), defaultdict(int)
for row in csv.reader(raw.splitlines()):
    if not row: continue
    if row[0] == "Kernel Name":
        if kern is not None and ex: break   # first matching kernel only
        kern = row[1] if (pat is None or pat.search(row[1])) else None
        continue
    if row[0] == "Address": hdr = row; continue
    if kern is None or hdr is None: continue
    op = row[hdr.index("Source")].split()
    op = op[1] if op[0].startswith("@") else op[0]
    ex[op] += int(row[hdr.index("Instructions Executed")]); st[op] += int(row[hdr.index("# Samples")])
tot, tots = sum(ex.values()), sum(st.values()) or 1
print(kern, "executed warp instructions:", tot)
for op, n in sorted(ex.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{op:28s} {n:12d} {100*n/tot:5.1f}%   samples {100*st[op]/tots:5.1f}%")
