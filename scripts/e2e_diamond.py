"""End to end through the frame API, 1920x1080 YUV420P16 pinned host frames, T host threads:
  flt = src.vszip.BoxBlur(hradius=2, vradius=2); out = flt.vszip.LimitFilter(src, dark_thr=4, bright_thr=4, elast=2)
evaluated as two get_frame calls (3 uploads + 2 downloads of 6.2 MB per frame) against one fused vszip_chain_get_frame with the
LimitFilter as a "diamond" element reading the chain's source (1 upload + 1 download), and the reference's AdaptiveBinarize usage
  out = src.vszip.AdaptiveBinarize(src.vszip.BoxBlur(hradius=5, vradius=5)) on YUV420P8 the same two ways.
usage: python scripts/e2e_diamond.py [frames] [threads]"""
import ctypes as C
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import vapoursynth_zip_b200 as vz

NE = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 8
W, H = 1920, 1080
vz.core.init([0])
lib = vz.load_library()
keep = []


def pinned_frame(dtype, fill):
    it = np.dtype(dtype).itemsize
    n = W * H + 2 * (W // 2) * (H // 2)
    t = torch.empty(n * it, dtype=torch.uint8).pin_memory()
    a = t.numpy().view(dtype)
    if fill:
        a[:] = np.random.default_rng(len(keep)).integers(0, 256 ** it, size=n, dtype=np.uint32).astype(dtype)
    keep.append(t)
    y = a[:W * H].reshape(H, W)
    u = a[W * H:W * H + (W // 2) * (H // 2)].reshape(H // 2, W // 2)
    v = a[W * H + (W // 2) * (H // 2):].reshape(H // 2, W // 2)
    return [y, u, v]


def run(fn, reps=3):
    with ThreadPoolExecutor(NT) as ex:
        list(ex.map(fn, range(NE)))
        t0 = time.perf_counter()
        for _ in range(reps):
            list(ex.map(fn, range(NE)))
        return reps * NE / (time.perf_counter() - t0)


def case(name, fmt, dtype, first, second, call_second):
    src = [pinned_frame(dtype, True) for _ in range(NE)]
    mid = [pinned_frame(dtype, False) for _ in range(NE)]
    dst = [pinned_frame(dtype, False) for _ in range(NE)]
    fs, fm, fd = ([vz._cframe(p) for p in fr] for fr in (src, mid, dst))
    handles = (C.c_void_p * 2)(first.handle, second.handle)
    chain = lib.vszip_chain_create(handles, 2)
    assert chain, vz._last_error()

    def unfused(i):
        assert lib.vszip_boxblur_get_frame(first.handle, i, C.byref(fs[i]), C.byref(fm[i])) == 0, vz._last_error()
        assert call_second(i, fs[i], fm[i], fd[i]) == 0, vz._last_error()

    def fused(i):
        assert lib.vszip_chain_get_frame(chain, i, C.byref(fs[i]), C.byref(fd[i]), (C.c_void_p * 2)(None, None)) == 0, vz._last_error()

    fa = run(unfused)
    ref_out = [[p.copy() for p in d] for d in dst[:2]]
    for d in dst:
        for p in d:
            p[:] = 0
    fb = run(fused)
    same = all(np.array_equal(a, b) for i in range(2) for a, b in zip(dst[i], ref_out[i]))
    fbytes = sum(p.nbytes for p in src[0])
    print(f"{name} {W}x{H} {fmt}, {NT} host threads, pinned frames: two get_frame calls {fa:.0f} fps, fused chain {fb:.0f} fps ({fb / fa:.2f}x), "
          f"outputs identical: {same}; PCIe bytes per frame: {5 * fbytes / 1e6:.1f} MB vs {2 * fbytes / 1e6:.1f} MB")
    lib.vszip_chain_free(chain)


vi16 = vz._vi(vz.FORMATS["YUV420P16"], W, H, 5000)
blur = vz.BoxBlurFilter(vi16, hradius=2, vradius=2)
lf = vz.LimitFilterFilter(vi16, vi16, None, dark_thr=4, bright_thr=4, elast=2)
case("BoxBlur(2,2) -> LimitFilter(flt, src)", "YUV420P16", np.uint16, blur, lf,
     lambda i, s, m, d: lib.vszip_limitfilter_get_frame(lf.handle, i, C.byref(m), C.byref(s), None, C.byref(d)))
vi8 = vz._vi(vz.FORMATS["YUV420P8"], W, H, 5000)
blur8 = vz.BoxBlurFilter(vi8, hradius=5, vradius=5)
ab = vz.AdaptiveBinarizeFilter(vi8, vi8, c=3)
case("AdaptiveBinarize(src, BoxBlur(src,5,5))", "YUV420P8", np.uint8, blur8, ab,
     lambda i, s, m, d: lib.vszip_adaptivebinarize_get_frame(ab.handle, i, C.byref(s), C.byref(m), C.byref(d)))
