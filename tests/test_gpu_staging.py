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
