"""End-to-end probe: frames/s through vszip_boxblur_get_frame for different numbers of host threads."""
import ctypes as C, sys, time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import vapoursynth_zip_b200 as vz
W, H = 1920, 1080
FB = (W * H + 2 * (W // 2) * (H // 2)) * 2
vz.core.init([0])
lib = vz.load_library()
flt = vz.BoxBlurFilter(vz._vi(vz.FORMATS["YUV420P16"], W, H, 1000), hradius=13, hpasses=5, vradius=13, vpasses=5)
ne = 64
def planes_of(t):
    a = t.numpy().view(np.uint16); out = []; off = 0
    for (h, w) in [(H, W), (H // 2, W // 2), (H // 2, W // 2)]:
        out.append(a[off:off + h * w].reshape(h, w)); off += h * w
    return out
hin = [torch.randint(0, 255, (FB,), dtype=torch.uint8).pin_memory() for _ in range(ne)]
hout = [torch.empty(FB, dtype=torch.uint8).pin_memory() for _ in range(ne)]
fi = [vz._cframe(planes_of(t)) for t in hin]; fo = [vz._cframe(planes_of(t)) for t in hout]
def one(i):
    rc = lib.vszip_boxblur_get_frame(flt.handle, i, C.byref(fi[i]), C.byref(fo[i]))
    assert rc == 0, vz._last_error()
for nt in (1, 2, 4, 8, 12, 16):
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(one, range(ne)))
        t0 = time.perf_counter()
        for _ in range(4): list(ex.map(one, range(ne)))
        dt = time.perf_counter() - t0
    print(nt, "threads:", round(4 * ne / dt), "fps", round(4 * ne * 2 * FB / dt / 1e9, 1), "GB/s both directions")
