"""Turns one GPU visit's raw output (gpurun_out/*_<round>.*) into the tracked evidence under profiles/.
usage: python scripts/make_profiles.py r01"""
import csv
import json
import shutil
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
P.mkdir(exist_ok=True)

for name in (f"bench_{R}.json", f"bench_ref_{R}.json", f"bench_extras_{R}.json", f"bench_extras_{R}.txt", f"pytest_gpu_{R}.log", f"e2e_chain_{R}.txt", f"e2e_diamond_{R}.txt", f"smoke_{R}.log"):
    if (G / name).exists():
        shutil.copy(G / name, P / name)

# ---- launch list of `bench.py --steps 2 --warmup 1` (ncu --metrics gpu__time_duration.sum): per-kernel shares
lf = G / f"launches_{R}.csv"
if lf.exists():
    rows = [r for r in csv.reader(lf.read_text().splitlines()) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        k = r[4].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e3
    tot = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list, `python bench.py --steps 2 --warmup 1 --no-cpu` ({R})", "",
             "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
             "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
    (P / f"launches_{R}.md").write_text("\n".join(lines) + "\n")
    shutil.copy(lf, P / f"launches_{R}.csv")

# ---- ncu --set full summaries
# (gpu_round.sh already summarises the captures on the box; a .ncu-rep that travelled back is summarised here)
for md in sorted(G.glob(f"ncu_*_{R}.md")):
    (P / md.name).write_text(f"# {md.stem}.ncu-rep (ncu --set full --clock-control none)\n\n" + md.read_text())
for rep in sorted(G.glob(f"ncu_*_{R}.ncu-rep")):
    if (G / (rep.stem + ".md")).exists():
        continue
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_summary.py"), str(rep)], capture_output=True, text=True).stdout
    (P / (rep.stem + ".md")).write_text(f"# {rep.name} (ncu --set full --clock-control none)\n\n" + out)
# ---- DRAM bytes per frame of the headline kernels (bench.py reads this for roofline.traffic): from the boxblur capture, 128 frames per launch
md = P / f"ncu_boxblur_{R}.md"
if md.exists():
    import re
    frames, per, kern = 128, {}, None
    for line in md.read_text().splitlines():
        m = re.match(r"### void .*?(\w+_kernel)<", line)
        if m:
            kern = m.group(1); per.setdefault(kern, 0.0); continue
        m = re.match(r"- dram__bytes_(read|write)\.sum: ([0-9.]+) (\w+)", line)
        if m and kern:
            per[kern] += float(m.group(2)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(3)]
    (P / f"traffic_{R}.json").write_text(json.dumps({"source": f"profiles/ncu_boxblur_{R}.md ({frames} frames per launch)",
                                                     "dram_bytes_per_frame": {k: v / frames for k, v in per.items()}}, indent=1) + "\n")
for name in (f"sanitizer_memcheck_{R}.log", f"sanitizer_racecheck_{R}.log", f"sanitizer_synccheck_{R}.log"):
    if (G / name).exists():
        shutil.copy(G / name, P / name)
print("profiles updated:", sorted(p.name for p in P.iterdir()))
