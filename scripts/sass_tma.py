"""Writes profiles/sass_tma_rNN.txt: per kernel family the static counts of the TMA / mbarrier instructions in the built library.
usage: python scripts/sass_tma.py r02"""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = subprocess.run(["cuobjdump", "-sass", str(ROOT / "vapoursynth_zip_b200/lib/libvszip_cuda.so")], capture_output=True, text=True).stdout
fn, counts = None, collections.OrderedDict()
pat = re.compile(r"\b(UBLKCP|UTMALDG|UTMASTG|SYNCS|LDGSTS|ELECT)\b")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); counts[fn] = collections.Counter(); continue
    if fn:
        m = pat.search(line)
        if m:
            counts[fn][m.group(1)] += 1
agg, tot = collections.OrderedDict(), collections.Counter()
for name, c in counts.items():
    tot += c
    if c["UBLKCP"] or c["UTMALDG"] or c["UTMASTG"]:
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        key = re.sub(r"<.*", "", re.sub(r"^void |vsz::|\(anonymous namespace\)::", "", d))
        a = agg.setdefault(key, {"n": 0, "c": collections.Counter()})
        a["n"] += 1; a["c"] += c
lines = [f"# SASS evidence for the TMA / mbarrier paths (cuobjdump -sass vapoursynth_zip_b200/lib/libvszip_cuda.so, sm_100a), round {rnd}",
         "# UBLKCP = cp.async.bulk (1-D bulk copy), UTMALDG = cp.async.bulk.tensor load (tensor-map tile), SYNCS = mbarrier operations,",
         "# LDGSTS = cp.async (per-thread, the round-1 streaming kernels).  Static instruction counts summed over the instantiations (radii 1..22).",
         f"# whole library: {len(counts)} kernels; UBLKCP {tot['UBLKCP']}, UTMALDG {tot['UTMALDG']}, UTMASTG {tot['UTMASTG']}, SYNCS {tot['SYNCS']}, LDGSTS {tot['LDGSTS']}",
         "", f"{'kernel':28s} {'instantiations':>14s} {'UBLKCP':>7s} {'UTMALDG':>8s} {'SYNCS':>6s} {'ELECT':>6s}"]
for k, a in agg.items():
    c = a["c"]
    lines.append(f"{k:28s} {a['n']:14d} {c['UBLKCP']:7d} {c['UTMALDG']:8d} {c['SYNCS']:6d} {c['ELECT']:6d}")
(ROOT / "profiles" / f"sass_tma_{rnd}.txt").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
