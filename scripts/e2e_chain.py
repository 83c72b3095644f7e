"""End-to-end BASELINE config 5 chain (BoxBlur(13,1,13,1) -> Bilateral(2,2) -> PlaneMinMax(0.1,0.1,planes=[0])) on
3840x2160 YUV444PS host frames through the frame API: three separate get_frame calls (3 uploads, 2 downloads per
frame) against one fused vszip_chain_get_frame (1 upload, 1 download).  Pinned host frames, T host threads.
usage: python scripts/e2e_chain.py [frames] [threads] [w] [h]"""
import ctypes as C
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import vapoursynth_zip_b200 as vz

NE = int(sys.argv[1]) if len(sys.argv) > 1 else 8
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W = int(sys.argv[3]) if len(sys.argv) > 3 else 3840
H = int(sys.argv[4]) if len(sys.argv) > 4 else 2160
FMT = "YUV444PS"
PB = W * H * 4
vz.core.init([0])
lib = vz.load_library()
vi = vz._vi(vz.FORMATS[FMT], W, H, 5000)
blur = vz.BoxBlurFilter(vi, hradius=13, hpasses=1, vradius=13, vpasses=1)
bil = vz.BilateralFilter(vi, sigmaS=2, sigmaR=2)
mm = vz.PlaneMinMaxFilter(vi, minthr=0.1, maxthr=0.1, planes=[0])
handles = (C.c_void_p * 3)(blur.handle, bil.handle, mm.handle)
chain = lib.vszip_chain_create(handles, 3)
assert chain, vz._last_error()


def pinned_frame(fill):
    t = torch.empty(3 * PB, dtype=torch.uint8).pin_memory()
    a = t.numpy().view(np.float32).reshape(3, H, W)
    if fill:
        rng = np.random.default_rng(len(keep))
        a[0] = rng.random((H, W), dtype=np.float32)
        a[1:] = rng.random((2, H, W), dtype=np.float32) - 0.5
    keep.append(t)
    return [a[0], a[1], a[2]]


keep = []
src = [pinned_frame(True) for _ in range(NE)]
mid = [pinned_frame(False) for _ in range(NE)]
dst = [pinned_frame(False) for _ in range(NE)]
fs, fm, fd = ([vz._cframe(p) for p in fr] for fr in (src, mid, dst))
res_a, res_b = [None] * NE, [None] * NE


def unfused(i):
    assert lib.vszip_boxblur_get_frame(blur.handle, i, C.byref(fs[i]), C.byref(fm[i])) == 0, vz._last_error()
    assert lib.vszip_bilateral_get_frame(bil.handle, i, C.byref(fm[i]), None, C.byref(fd[i])) == 0, vz._last_error()
    o = vz._MinMaxProps()
    assert lib.vszip_planeminmax_get_frame(mm.handle, i, C.byref(fd[i]), None, C.byref(o)) == 0, vz._last_error()
    res_a[i] = (o.fmin[0], o.fmax[0])


def fused(i):
    o = vz._MinMaxProps()
    outs = (C.c_void_p * 3)(None, None, C.cast(C.pointer(o), C.c_void_p))
    assert lib.vszip_chain_get_frame(chain, i, C.byref(fs[i]), C.byref(fd[i]), outs) == 0, vz._last_error()
    res_b[i] = (o.fmin[0], o.fmax[0])


def run(fn, reps=3):
    with ThreadPoolExecutor(NT) as ex:
        list(ex.map(fn, range(NE)))
        t0 = time.perf_counter()
        for _ in range(reps):
            list(ex.map(fn, range(NE)))
        return reps * NE / (time.perf_counter() - t0)


fa = run(unfused)
ref_out = [np.stack(d).copy() for d in dst[:2]]
fb = run(fused)
same = all(np.array_equal(np.stack(dst[i]), ref_out[i]) for i in range(2)) and res_a == res_b
print(f"config-5 chain {W}x{H} {FMT}, {NT} host threads, pinned frames: unfused {fa:.1f} fps, fused {fb:.1f} fps "
      f"({fb / fa:.2f}x), outputs identical: {same}; bytes per frame over PCIe: unfused {5 * 3 * PB / 1e6:.0f} MB, fused {2 * 3 * PB / 1e6:.0f} MB")
