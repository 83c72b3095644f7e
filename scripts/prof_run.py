"""Small driver for ncu captures: runs one filter over a device-resident noise batch a few times.
usage: python scripts/prof_run.py boxblur|boxblur_ct|bilateral|pbfic|limitfilter|binarize|planestats|minmax|average [frames] [reps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import vapoursynth_zip_b200 as vz

what = sys.argv[1] if len(sys.argv) > 1 else "boxblur"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
vz.core.init([0])
if what in ("minmax", "average", "planestats"):
    fmt, w, h = "GRAY16", 3840, 2160
elif what in ("boxblur_ctf", "bilateral_f32"):
    fmt, w, h = "YUV444PS", 3840, 2160
else:
    fmt, w, h = "YUV420P16", 1920, 1080
src = vz.DeviceClip(fmt, w, h, frames)
dst = vz.DeviceClip(fmt, w, h, frames)
src.fill_noise(seed=1234)
if what == "boxblur":
    f = vz.BoxBlurFilter(src.info(), hradius=13, hpasses=5, vradius=13, vpasses=5)
    run = lambda: f.run_device(src, dst)
elif what == "boxblur_h1":
    f = vz.BoxBlurFilter(src.info(), hradius=13, hpasses=1, vradius=0, vpasses=0)
    run = lambda: f.run_device(src, dst)
elif what == "boxblur_v1":
    f = vz.BoxBlurFilter(src.info(), hradius=0, hpasses=0, vradius=13, vpasses=1)
    run = lambda: f.run_device(src, dst)
elif what == "boxblur_ct":
    f = vz.BoxBlurFilter(src.info(), hradius=13, hpasses=1, vradius=13, vpasses=1)
    run = lambda: f.run_device(src, dst)
elif what == "boxblur_ctf":   # config 5a: comptime float path
    f = vz.BoxBlurFilter(src.info(), hradius=13, hpasses=1, vradius=13, vpasses=1)
    run = lambda: f.run_device(src, dst)
elif what == "bilateral_f32":  # config 5b
    f = vz.BilateralFilter(src.info(), sigmaS=2, sigmaR=2)
    run = lambda: f.run_device(src, dst)
elif what == "bilateral":
    f = vz.BilateralFilter(src.info(), sigmaS=2, sigmaR=2)
    run = lambda: f.run_device(src, dst)
elif what == "pbfic":
    f = vz.BilateralFilter(src.info(), sigmaS=8, sigmaR=0.1, planes=[0])
    run = lambda: f.run_device(src, dst)
elif what == "limitfilter":   # image-like content: src = smoothed noise, flt = BoxBlur(src, 2), ref = BoxBlur(src, 4)
    tmp, flt, ref = (vz.DeviceClip(fmt, w, h, frames) for _ in range(3))
    vz.BoxBlurFilter(src.info(), hradius=13, hpasses=3, vradius=13, vpasses=3).run_device(src, tmp)
    vz.BoxBlurFilter(src.info(), hradius=2, vradius=2).run_device(tmp, flt)
    vz.BoxBlurFilter(src.info(), hradius=4, vradius=4).run_device(tmp, ref)
    f = vz.LimitFilterFilter(src.info(), src.info(), src.info(), dark_thr=1, bright_thr=1, elast=2)
    run = lambda: f.run_device(flt, tmp, dst, ref=ref)
elif what == "binarize":
    src.free(); dst.free()
    src, dst = vz.DeviceClip("YUV420P8", w, h, frames), vz.DeviceClip("YUV420P8", w, h, frames)
    b2 = vz.DeviceClip("YUV420P8", w, h, frames)
    src.fill_noise(seed=1234); b2.fill_noise(seed=99)
    f = vz.AdaptiveBinarizeFilter(src.info(), src.info(), c=3)
    run = lambda: f.run_device(src, b2, dst)
elif what == "planestats":   # PlaneMinMax(thr) + PlaneAverage from one read (fused bracket kernel)
    mmf = vz.PlaneMinMaxFilter(src.info(), minthr=0.1, maxthr=0.1)
    avf = vz.PlaneAverageFilter(src.info(), exclude=[0, 32768])
    run = lambda: vz.plane_stats_device(mmf, avf, src, fetch=False)
elif what == "minmax":
    f = vz.PlaneMinMaxFilter(src.info(), minthr=0.1, maxthr=0.1)
    run = lambda: f.run_device(src)
else:
    f = vz.PlaneAverageFilter(src.info(), exclude=[0, 32768])
    run = lambda: f.run_device(src)
for _ in range(reps):
    run()
vz.core.sync()
print("done", what, frames, reps, vz.core.kernel_launches)
