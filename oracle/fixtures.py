"""Rebuilds the reference test-suite clips (tests/conftest.py:72-121 of the reference) from the
committed 640x320 RGB crop, using oracle/fixture.cpp for the zimg-equivalent arithmetic.

TEST INFRASTRUCTURE ONLY.  A clip here is just {"format": name, "planes": [np.ndarray, ...]}.
"""
import ctypes as C
from functools import lru_cache
from pathlib import Path

import numpy as np
from PIL import Image

from . import lib

GOLDEN_DIR = Path(__file__).resolve().parents[1] / "tests" / "golden"

# name -> (family, sample 'i'/'f', bits, ssw, ssh)
FORMATS = {
    "GRAY8": ("GRAY", "i", 8, 0, 0), "GRAY10": ("GRAY", "i", 10, 0, 0), "GRAY16": ("GRAY", "i", 16, 0, 0),
    "GRAYH": ("GRAY", "f", 16, 0, 0), "GRAYS": ("GRAY", "f", 32, 0, 0), "GRAY32": ("GRAY", "i", 32, 0, 0),
    "YUV420P8": ("YUV", "i", 8, 1, 1), "YUV420P10": ("YUV", "i", 10, 1, 1), "YUV420P16": ("YUV", "i", 16, 1, 1),
    "YUV420PS": ("YUV", "f", 32, 1, 1),
    "YUV444P8": ("YUV", "i", 8, 0, 0), "YUV444P16": ("YUV", "i", 16, 0, 0), "YUV444PS": ("YUV", "f", 32, 0, 0),
    "RGB24": ("RGB", "i", 8, 0, 0), "RGB48": ("RGB", "i", 16, 0, 0),
    "RGBH": ("RGB", "f", 16, 0, 0), "RGBS": ("RGB", "f", 32, 0, 0),
}


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


@lru_cache(maxsize=None)
def src_rgb24():
    im = np.asarray(Image.open(GOLDEN_DIR / "src_rgb_640x320.png").convert("RGB"))
    assert im.shape == (320, 640, 3)
    return [np.ascontiguousarray(im[:, :, c]) for c in range(3)]


@lru_cache(maxsize=None)
def _rgbs():
    out = []
    for p in src_rgb24():
        o = np.empty(p.shape, np.float32)
        lib().vsf_rgb8_to_rgbs(_vp(p), p.size, _vp(o))
        out.append(o)
    return out


@lru_cache(maxsize=None)
def _yuv_full(plane: int):
    r, g, b = _rgbs()
    o = np.empty(r.shape, np.float32)
    lib().vsf_rgbs_to_yuv_plane(_vp(r), _vp(g), _vp(b), r.size, plane, _vp(o))
    return o


def _sub420(p):
    h, w = p.shape
    o = np.empty((h // 2, w // 2), np.float32)
    lib().vsf_chroma_420(_vp(np.ascontiguousarray(p)), w, h, _vp(o))
    return o


def _quant(p, bits, chroma):
    o = np.empty(p.shape, np.uint16)
    lib().vsf_quantise(_vp(np.ascontiguousarray(p)), p.size, bits, int(chroma), _vp(o))
    return o.astype(np.uint8) if bits == 8 else o


def _to_f16(p):
    o = np.empty(p.shape, np.uint16)
    lib().vsf_f32_to_f16(_vp(np.ascontiguousarray(p)), p.size, _vp(o))
    return o.view(np.float16)


@lru_cache(maxsize=None)
def _full(fmt: str):
    fam, st, bits, ssw, ssh = FORMATS[fmt]
    if fam == "RGB":
        if st == "i":
            if bits == 8:
                return [p.copy() for p in src_rgb24()]
            assert bits == 16, "only RGB24/RGB48 are used by the golden keys"
            return [(p.astype(np.uint16) * 257) for p in src_rgb24()]
        return [p.copy() for p in _rgbs()] if bits == 32 else [_to_f16(p) for p in _rgbs()]
    nplanes = 1 if fam == "GRAY" else 3
    planes = []
    for pl in range(nplanes):
        p = _yuv_full(pl)
        if pl > 0 and ssw == 1 and ssh == 1:
            p = _sub420(p)
        if st == "i":
            planes.append(_quant(p, bits, pl > 0))
        elif bits == 32:
            planes.append(p.copy())
        else:
            planes.append(_to_f16(p))
    return planes


def make_clip(fmt: str, geometry: str = "full"):
    """The reference's make_clip(fmt, geometry) fixture (tests/conftest.py:108-139)."""
    fam, st, bits, ssw, ssh = FORMATS[fmt]
    planes = _full(fmt)
    wmod, hmod = 1 << ssw, 1 << ssh

    def cut(pl, p, x0, y0, w, h):
        sx, sy = (ssw, ssh) if pl > 0 else (0, 0)
        return np.ascontiguousarray(p[y0 >> sy:(y0 + h) >> sy, x0 >> sx:(x0 + w) >> sx])

    H, W = planes[0].shape
    if geometry == "full":
        out = [np.ascontiguousarray(p) for p in planes]
    elif geometry == "odd":
        out = [cut(i, p, 0, 0, W - wmod, H - hmod) for i, p in enumerate(planes)]
    elif geometry == "tiny":
        out = [cut(i, p, 200, 100, 13 - 13 % wmod, 7 - 7 % hmod) for i, p in enumerate(planes)]
    else:
        raise ValueError(geometry)
    return {"format": fmt, "planes": out}
