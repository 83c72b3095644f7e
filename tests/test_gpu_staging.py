"""Host <-> device staging variants of get_frame: pageable planes (staged through the slot's pinned buffer), pinned
planes with the device pitch (one linear DMA per plane, or one per frame when the planes are contiguous), pinned
planes with padded strides (2-D DMA).  All must give the same result."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, noise_clip

pytestmark = pytest.mark.gpu


def _pinned_planes(shapes, dtype, layout):
    item = np.dtype(dtype).itemsize
    if layout == "contiguous":   # one buffer, planes back to back, stride == row bytes
        total = sum(h * w * item for h, w in shapes)
        buf = torch.empty(total, dtype=torch.uint8).pin_memory()
        flat, out, off = buf.numpy(), [], 0
        for h, w in shapes:
            out.append(flat[off:off + h * w * item].view(dtype).reshape(h, w))
            off += h * w * item
        return out, buf
    keep, out = [], []
    for h, w in shapes:           # separate buffers; "padded": rows 24 samples longer than the plane
        pw = w + (24 if layout == "padded" else 0)
        buf = torch.empty(h * pw * item, dtype=torch.uint8).pin_memory()
        keep.append(buf)
        out.append(buf.numpy().view(dtype).reshape(h, pw)[:, :w])
    return out, keep


@pytest.mark.parametrize("layout", ["contiguous", "separate", "padded"])
@pytest.mark.parametrize(("fmt", "w", "h"), [("YUV420P16", 1920, 1080), ("YUV420P16", 322, 182), ("GRAYS", 517, 243), ("YUV444P8", 640, 360)])
def test_pinned_frames(fmt, w, h, layout):
    clip = noise_clip(fmt, w, h, seed=33)
    shapes = [p.shape for p in clip["planes"]]
    src, k1 = _pinned_planes(shapes, clip["planes"][0].dtype, layout)
    dst, k2 = _pinned_planes(shapes, clip["planes"][0].dtype, layout)
    for s, p in zip(src, clip["planes"]):
        s[...] = p
    vz.core._ensure_init()
    f = vz.BoxBlurFilter(vz._vi(vz.FORMATS[fmt], w, h, 1), hradius=3, hpasses=2, vradius=2, vpasses=1)
    fs, fd = vz._cframe(src), vz._cframe(dst)
    assert vz.load_library().vszip_boxblur_get_frame(f.handle, 0, C.byref(fs), C.byref(fd)) == 0, vz._last_error()
    want = oa.boxblur(clip, hradius=3, hpasses=2, vradius=2, vpasses=1)
    assert_same_planes([np.ascontiguousarray(d) for d in dst], want["planes"], f"pinned {layout} {fmt}")


def test_pageable_buffers_are_registered_on_second_sight_and_forgotten():
    """The host pin cache (csrc/runtime.cu): a pageable plane buffer that comes back is page-locked and DMA'd in place; results
    stay identical whichever path a call took, and vszip_cuda_host_forget releases the registration."""
    lib = vz.load_library()
    vz.core._ensure_init()
    lib.vszip_cuda_host_forget(None)
    assert lib.vszip_cuda_host_registered_bytes() == 0
    fmt, w, h = "YUV420P16", 640, 360
    clip = noise_clip(fmt, w, h, seed=5)
    src = [np.array(p, copy=True) for p in clip["planes"]]       # plain pageable numpy memory, one allocation per plane
    dst = [np.zeros_like(p) for p in src]
    f = vz.BoxBlurFilter(vz._vi(vz.FORMATS[fmt], w, h, 1), hradius=4, hpasses=2, vradius=3, vpasses=2)
    want = oa.boxblur(clip, hradius=4, hpasses=2, vradius=3, vpasses=2)["planes"]
    fs, fd = vz._cframe(src), vz._cframe(dst)
    seen = []
    for call in range(4):
        for d in dst:
            d[...] = 0
        assert lib.vszip_boxblur_get_frame(f.handle, call, C.byref(fs), C.byref(fd)) == 0, vz._last_error()
        assert_same_planes(dst, want, f"call {call}")
        seen.append(lib.vszip_cuda_host_registered_bytes())
    assert seen[0] == 0                                            # first sighting: staged copy
    assert seen[1] >= sum(p.nbytes for p in src + dst)             # second sighting: registered (page-rounded)
    assert seen[2] == seen[1] == seen[3]                           # and kept, not re-registered
    lib.vszip_cuda_host_forget(C.c_void_p(src[0].ctypes.data))
    assert lib.vszip_cuda_host_registered_bytes() < seen[1]
    lib.vszip_cuda_host_forget(None)
    assert lib.vszip_cuda_host_registered_bytes() == 0
    assert lib.vszip_boxblur_get_frame(f.handle, 9, C.byref(fs), C.byref(fd)) == 0   # back on the staging path, still correct
    assert_same_planes(dst, want, "after forget")
