"""H-only / V-only split of the comptime float BoxBlur is not reachable through the API (the comptime path always runs both), so this
times the whole filter; used with the diagnostic builds of boxblur_ctf.cu (VSZ_CTF_DIAG_*)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vapoursynth_zip_b200 as vz
vz.core.init([0])
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
src = vz.DeviceClip("YUV444PS", 3840, 2160, 8); dst = vz.DeviceClip("YUV444PS", 3840, 2160, 8)
src.fill_noise(1234)
f = vz.BoxBlurFilter(src.info(), hradius=13, vradius=13)
for _ in range(2): f.run_device(src, dst, 0, 8, st.cuda_stream)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): f.run_device(src, dst, 0, 8, st.cuda_stream)
b.record(); torch.cuda.synchronize()
print("us/frame %.1f" % (a.elapsed_time(b) / 5 * 1000 / 8))
