"""PlaneMinMax(0.1, 0.1) on 4K GRAY16 content that stresses the sampled single-read path: us/frame per content type."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import vapoursynth_zip_b200 as vz

W, H, N = 3840, 2160, 16
vz.core.init([0])
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
rng = np.random.default_rng(1)
y, x = np.mgrid[0:H, 0:W]
contents = {
    "noise": lambda: rng.integers(0, 65536, (H, W)).astype(np.uint16),
    "constant": lambda: np.full((H, W), 12345, np.uint16),
    "dark+-2": lambda: (4096 + rng.integers(-2, 3, (H, W))).astype(np.uint16),
    "gradient": lambda: ((x * 65535) // (W - 1)).astype(np.uint16),
    "letterbox": lambda: np.where((y < H // 6) | (y >= H - H // 6), 4096, rng.integers(8000, 60000, (H, W))).astype(np.uint16),
    "8bit<<8": lambda: (rng.integers(0, 256, (H, W)) << 8).astype(np.uint16),
}
f = None
for name, make in contents.items():
    clip = vz.DeviceClip("GRAY16", W, H, N)
    plane = make()
    for i in range(N):
        clip.upload(i, [plane])
    f = vz.PlaneMinMaxFilter(clip.info(), minthr=0.1, maxthr=0.1)
    for _ in range(2):
        f.run_device(clip, stream=st.cuda_stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        r = f.run_device(clip, stream=st.cuda_stream)
    b.record(); torch.cuda.synchronize()
    print(f"{name:10s} {a.elapsed_time(b) / 5 / N * 1e3:8.2f} us/frame  {r[0]}")
    clip.free()
