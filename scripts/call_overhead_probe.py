"""How many get_frame calls per second does one process sustain on frames too small for PCIe or the kernels to matter?
(64x64 GRAY8 pinned frames through vszip_limiter_get_frame: 1 upload, 1 launch, 1 download, 1 sync per call.)
usage: python scripts/call_overhead_probe.py"""
import ctypes as C
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import vapoursynth_zip_b200 as vz

vz.core.init([0])
lib = vz.load_library()
W = H = 64
vi = vz._vi(vz.FORMATS["GRAY8"], W, H, 100000)
f = vz.LimiterFilter(vi, tv_range=True)
keep = []


def frame():
    t = torch.zeros(W * H, dtype=torch.uint8).pin_memory()
    keep.append(t)
    return vz._cframe([t.numpy().reshape(H, W)])


N = 64
src, dst = [frame() for _ in range(N)], [frame() for _ in range(N)]


def one(i):
    assert lib.vszip_limiter_get_frame(f.handle, i, C.byref(src[i % N]), C.byref(dst[i % N])) == 0


for nt in (1, 2, 4, 8, 16):
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(one, range(256)))
        t0 = time.perf_counter()
        list(ex.map(one, range(4096)))
        dt = time.perf_counter() - t0
    print(f"{nt:2d} host threads: {4096 / dt:8.0f} calls/s  ({dt / 4096 * 1e6:6.1f} us per call across the process)")
