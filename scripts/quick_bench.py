"""Device-only timing of one filter configuration (for A/B experiments).
usage: python scripts/quick_bench.py [frames] [reps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vapoursynth_zip_b200 as vz

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
vz.core.init([0])
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
src = vz.DeviceClip("YUV420P16", 1920, 1080, frames); dst = vz.DeviceClip("YUV420P16", 1920, 1080, frames)
src.fill_noise(1234)
def t(f):
    for _ in range(2): f.run_device(src, dst, 0, frames, st.cuda_stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f.run_device(src, dst, 0, frames, st.cuda_stream)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
cfgs = {"H5": dict(hradius=13, hpasses=5, vradius=0, vpasses=0), "V5": dict(hradius=0, hpasses=0, vradius=13, vpasses=5),
        "HV5": dict(hradius=13, hpasses=5, vradius=13, vpasses=5), "CT13": dict(hradius=13, vradius=13),
        "H1": dict(hradius=13, hpasses=1, vradius=0, vpasses=0), "V1": dict(hradius=0, hpasses=0, vradius=13, vpasses=1)}
for p in (2, 3, 4):
    cfgs["H%d" % p] = dict(hradius=13, hpasses=p, vradius=0, vpasses=0)
    cfgs["V%d" % p] = dict(hradius=0, hpasses=0, vradius=13, vpasses=p)
if len(sys.argv) > 3:
    cfgs = {k: cfgs[k] for k in sys.argv[3].split(",")}
out = {}
for k, a in cfgs.items():
    ms = t(vz.BoxBlurFilter(src.info(), **a))
    out[k] = round(ms * 1000 / frames, 2)
print("us/frame", out, ("fps(HV5)=%d" % (1e6 / out["HV5"])) if "HV5" in out else "")
