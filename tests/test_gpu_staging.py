"""Host <-> device staging variants of get_frame: pageable planes (staged through the slot's pinned buffer), pinned
planes with the device pitch (one linear DMA per plane, or one per frame when the planes are contiguous), pinned
planes with padded strides (2-D DMA).  All must give the same result."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, noise_clip, to_node

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
    """The opt-in host pin cache (csrc/runtime.cu): a pageable plane buffer that comes back is page-locked and DMA'd in place;
    results stay identical whichever path a call took, and vszip_cuda_host_forget releases the registration."""
    import mmap
    lib = vz.load_library()
    vz.core._ensure_init()
    lib.vszip_cuda_host_forget(None)
    assert lib.vszip_cuda_host_registered_bytes() == 0
    fmt, w, h = "YUV420P16", 640, 360
    clip = noise_clip(fmt, w, h, seed=5)
    maps = []

    def own_pages(p):   # one anonymous mapping per plane (what malloc returns for large buffers), alive until the test ends
        m = mmap.mmap(-1, p.nbytes + 4096)
        maps.append(m)
        a = np.frombuffer(m, dtype=p.dtype, count=p.size, offset=64).reshape(p.shape)
        a[...] = p
        return a
    src = [own_pages(p) for p in clip["planes"]]
    dst = [own_pages(np.zeros_like(p)) for p in clip["planes"]]
    f = vz.BoxBlurFilter(vz._vi(vz.FORMATS[fmt], w, h, 1), hradius=4, hpasses=2, vradius=3, vpasses=2)
    want = oa.boxblur(clip, hradius=4, hpasses=2, vradius=3, vpasses=2)["planes"]
    fs, fd = vz._cframe(src), vz._cframe(dst)

    def call(n, what):
        for d in dst:
            d[...] = 0
        assert lib.vszip_boxblur_get_frame(f.handle, n, C.byref(fs), C.byref(fd)) == 0, vz._last_error()
        assert_same_planes([np.ascontiguousarray(d) for d in dst], want, what)
    before = lib.vszip_cuda_host_register_limit(0)
    try:
        for n in range(3):
            call(n, f"default (off) call {n}")
        assert lib.vszip_cuda_host_registered_bytes() == 0            # off by default: nothing is ever registered
        lib.vszip_cuda_host_forget(None)
        lib.vszip_cuda_host_register_limit(1 << 30)
        seen = []
        for n in range(4):
            call(n, f"opt-in call {n}")
            seen.append(lib.vszip_cuda_host_registered_bytes())
        assert seen[0] == 0                                            # first sighting: staged copy
        assert seen[1] >= sum(p.nbytes for p in src + dst)             # second sighting: registered (page-rounded)
        assert seen[2] == seen[1] == seen[3]                           # and kept, not re-registered
        lib.vszip_cuda_host_forget(C.c_void_p(src[0].ctypes.data))
        assert lib.vszip_cuda_host_registered_bytes() < seen[1]
        lib.vszip_cuda_host_forget(None)
        assert lib.vszip_cuda_host_registered_bytes() == 0
        call(9, "after forget")                                        # back on the staging path, still correct
    finally:
        lib.vszip_cuda_host_forget(None)
        lib.vszip_cuda_host_register_limit(before)


def test_plane_that_only_starts_in_pinned_memory_is_staged():
    """A plane counts as application-pinned only if ALL of it lies inside one page-locked range: here the application
    registered the first half of the buffer only, so the copy must not be issued as a direct DMA."""
    import mmap
    import torch
    lib = vz.load_library()
    vz.core._ensure_init()
    w, h = 1024, 64                                                # 2 KiB rows: 32 rows = 16 pages
    clip = noise_clip("GRAY16", w, h, seed=8)
    want = oa.boxblur(clip, hradius=2, vradius=2)["planes"]
    m = mmap.mmap(-1, w * h * 2)
    src = np.frombuffer(m, dtype=np.uint16).reshape(h, w)
    src[...] = clip["planes"][0]
    rt = torch.cuda.cudart()
    assert int(rt.cudaHostRegister(src.ctypes.data, w * h, 0)) == 0   # first half of the plane
    try:
        dst = np.zeros((h, w), np.uint16)
        f = vz.BoxBlurFilter(vz._vi(vz.FORMATS["GRAY16"], w, h, 1), hradius=2, vradius=2)
        fs, fd = vz._cframe([src]), vz._cframe([dst])
        for n in range(2):
            assert lib.vszip_boxblur_get_frame(f.handle, n, C.byref(fs), C.byref(fd)) == 0, vz._last_error()
            assert_same_planes([dst], want, "half-registered plane")
    finally:
        lib.vszip_cuda_host_forget(None)
        rt.cudaHostUnregister(src.ctypes.data)


def test_reserve_sizes_the_slots_up_front():
    """vszip_cuda_reserve: the staging buffers of every slot exist before the first frame (no allocation on the request path); frames
    of that format and of a larger one still come out right (the buffers grow as before)."""
    small = noise_clip("YUV420P16", 320, 180, seed=5)
    node = to_node(small)
    vz.core.reserve(node._info(), buffers=3)
    got = node.vszip.BoxBlur(hradius=3, vradius=3).get_frame(0)
    assert_same_planes(got.planes, oa.boxblur(small, hradius=3, vradius=3)["planes"])
    big = noise_clip("YUV444P16", 640, 362, seed=6)
    got = to_node(big).vszip.BoxBlur(hradius=2, hpasses=2, vradius=2, vpasses=2).get_frame(0)
    assert_same_planes(got.planes, oa.boxblur(big, hradius=2, hpasses=2, vradius=2, vpasses=2)["planes"])
    with pytest.raises(vz.Error):
        bad = node._info()
        bad.width = 0
        vz.core.reserve(bad)
