"""Single-launch latency of BoxBlur / AdaptiveBinarize on 1 and 8 device-resident 1080p frames (the get_frame path launches one frame at a time).
usage: python scripts/latency_probe.py"""
import sys, torch
sys.path.insert(0, ".")
import vapoursynth_zip_b200 as vz
vz.core.init([0])
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def t(fn, reps=50):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000
for fmt in ("YUV420P8", "YUV420P16"):
    for frames in (1, 8):
        s, m, d = (vz.DeviceClip(fmt, 1920, 1080, frames) for _ in range(3))
        s.fill_noise(1)
        for name, r in (("BoxBlur(5,5)", dict(hradius=5, vradius=5)), ("BoxBlur(2,2)", dict(hradius=2, vradius=2)), ("BoxBlur(13,5,13,5)", dict(hradius=13, hpasses=5, vradius=13, vpasses=5))):
            f = vz.BoxBlurFilter(s.info(), **r)
            print(fmt, frames, "frames", name, "%.1f us per launch set" % t(lambda: f.run_device(s, m, stream=st.cuda_stream)))
        if fmt == "YUV420P8":
            ab = vz.AdaptiveBinarizeFilter(s.info(), s.info(), c=3)
            print(fmt, frames, "frames AdaptiveBinarize %.1f us" % t(lambda: ab.run_device(s, m, d, stream=st.cuda_stream)))
