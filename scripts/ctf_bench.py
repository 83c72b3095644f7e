"""Device-resident timing of the comptime float BoxBlur (config 5a) at a few batch sizes / radii / formats."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vapoursynth_zip_b200 as vz
vz.core.init([0])
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def t(fmt, w, h, frames, args, reps=5):
    src = vz.DeviceClip(fmt, w, h, frames); dst = vz.DeviceClip(fmt, w, h, frames)
    src.fill_noise(1234)
    f = vz.BoxBlurFilter(src.info(), **args)
    for _ in range(2): f.run_device(src, dst, 0, frames, st.cuda_stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f.run_device(src, dst, 0, frames, st.cuda_stream)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1000 / frames
    print(fmt, w, h, frames, args, "us/frame %.1f  algorithmic GB/s %.0f" % (us, 2 * src.frame_bytes / us / 1e3), flush=True)
    src.free(); dst.free()
for fr in (1, 8, 32):
    t("YUV444PS", 3840, 2160, fr, dict(hradius=13, vradius=13))
t("YUV444PS", 3840, 2160, 8, dict(hradius=13, hpasses=1, vradius=0, vpasses=0))
t("YUV444PS", 3840, 2160, 8, dict(hradius=4, vradius=4))
t("YUV444PS", 3840, 2160, 8, dict(hradius=22, vradius=22))
t("YUV444PH", 3840, 2160, 8, dict(hradius=13, vradius=13))
t("GRAYS", 1920, 1080, 64, dict(hradius=13, vradius=13))
