"""Per-call latency of PlaneMinMax(0.1, 0.1) on 1, 4 and 16 device-resident 4K frames (memsets + sampling + bracket + the two early-exit
fallback launches): ~66 us for a lone frame, of which the sampling kernel's dependent phases are ~30.
usage: python scripts/stats_latency_probe.py"""
import sys, torch
sys.path.insert(0, ".")
import vapoursynth_zip_b200 as vz
vz.core.init([0])
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
lib = vz.load_library()
def t(fn, reps=30):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000
for fmt in ("GRAY16", "GRAYS"):
    for frames in (1, 4, 16):
        s = vz.DeviceClip(fmt, 3840, 2160, frames); s.fill_noise(1)
        f = vz.PlaneMinMaxFilter(s.info(), minthr=0.1, maxthr=0.1)
        us = t(lambda: vz._check(lib.vszip_planeminmax_device(f.handle, s.handle, None, 0, frames, None, st.cuda_stream)))
        print(fmt, frames, "frames: %.1f us per call, %.2f us/frame" % (us, us / frames))
        s.free()
