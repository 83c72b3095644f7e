"""Regenerates tests/golden/ from the reference checkout (run in the build container only).

  python tests/golden/make_fixtures.py [/root/reference]

Outputs (committed, MIT-licensed test data from dnjulek/vapoursynth-zip):
  src_rgb_640x320.png   the reference suite's source fixture: tests/image.png cropped exactly
                        like tests/conftest.py:72-76 (left = width-640, bottom = height-320)
  boxblur.json, bilateral.json, planeminmax.json, planeaverage.json, limiter.json, limitfilter.json,
  adaptive_binarize.json
                        verbatim copies of tests/goldens/<name>.json (per-plane PlaneStats /
                        frame-prop snapshots recorded from the real Zig plugin)
The GPU box has no /root/reference, so the tests only ever read these copies.
"""
import json
import shutil
import sys
from pathlib import Path

from PIL import Image

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
out = Path(__file__).resolve().parent

im = Image.open(ref / "tests" / "image.png").convert("RGB")
w, h = im.size
crop = im.crop((w - 640, 0, w, 320))  # Crop(left=w-640, bottom=h-320) keeps rows 0:320, cols w-640:w
assert crop.size == (640, 320)
crop.save(out / "src_rgb_640x320.png", optimize=True)

for name in ("boxblur", "bilateral", "planeminmax", "planeaverage", "limiter", "limitfilter", "adaptive_binarize"):
    data = json.loads((ref / "tests" / "goldens" / f"{name}.json").read_text())
    (out / f"{name}.json").write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")
    print(name, len(data), "keys")
print("wrote", out)
