"""Small probe for the comptime float BoxBlur kernels: one call, checked against the oracle (debugging aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np
import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import noise_clip, to_node
fmt = sys.argv[1] if len(sys.argv) > 1 else "GRAYS"
w, h, r = (int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (331, 203, 2)))
clip = noise_clip(fmt, w, h, seed=1)
got = to_node(clip).vszip.BoxBlur(hradius=r, vradius=r).get_frame(0)
want = oa.boxblur(clip, hradius=r, vradius=r)
for i, (g, wv) in enumerate(zip(got.planes, want["planes"])):
    bad = np.argwhere(np.ascontiguousarray(g).view(np.uint32 if g.itemsize == 4 else np.uint16) != wv.view(np.uint32 if g.itemsize == 4 else np.uint16))
    print("plane", i, "mismatches", len(bad), "first", bad[:5].tolist())
