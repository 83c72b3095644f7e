"""vapoursynth_zip_b200 — Python host mirror of the vszip plugin surface for the B200 hot path.

The product is the C-ABI library (include/vszip_cuda.h, built into lib/libvszip_cuda.so from
csrc/*.cu).  This module is the thin host side used where the reference would be driven through
VapourSynth: it mirrors the `clip.vszip.BoxBlur / Bilateral / PlaneMinMax / PlaneAverage` call
surface (same names, arguments, defaults, frame props and error messages; src/vszip.zig:46-69,
:184-199) on top of a minimal clip/frame model, so the parity tests read like the reference's own
tests (tests/test_boxblur.py etc. of the reference).  All arithmetic happens in the CUDA library;
there is no CPU fallback and a missing library or GPU raises `Error`.
"""
from __future__ import annotations

import ctypes as C
import weakref
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

__all__ = ["Error", "core", "plane_stats_device", "VideoFormat", "VideoFrame", "VideoNode", "DeviceClip", "GRAY", "RGB", "YUV", "INTEGER", "FLOAT"]

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("VSZIP_CUDA_LIB") or (_PKG / "lib" / "libvszip_cuda.so"))  # override: A/B builds only

GRAY, RGB, YUV = 1, 2, 3          # VapourSynth4.h VSColorFamily
INTEGER, FLOAT = 0, 1             # VSSampleType


class Error(RuntimeError):
    """Mirror of vapoursynth.Error: raised with the reference's message on invalid arguments."""


# --------------------------------------------------------------------------- ctypes structs
class _VideoInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "num_frames", "color_family", "sample_type", "bits_per_sample",
                                         "bytes_per_sample", "sub_sampling_w", "sub_sampling_h", "num_planes")]


class _Frame(C.Structure):
    _fields_ = [("data", C.c_void_p * 3), ("stride", C.c_ssize_t * 3)]


class _BoxBlurArgs(C.Structure):
    _fields_ = [("planes", C.POINTER(C.c_int64)), ("num_planes", C.c_int32),
                ("has_hradius", C.c_int32), ("hradius", C.c_int64), ("has_hpasses", C.c_int32), ("hpasses", C.c_int64),
                ("has_vradius", C.c_int32), ("vradius", C.c_int64), ("has_vpasses", C.c_int32), ("vpasses", C.c_int64)]


class _BilateralArgs(C.Structure):
    _fields_ = [("sigmaS", C.POINTER(C.c_double)), ("num_sigmaS", C.c_int32), ("sigmaR", C.POINTER(C.c_double)), ("num_sigmaR", C.c_int32),
                ("planes", C.POINTER(C.c_int64)), ("num_planes", C.c_int32), ("algorithm", C.POINTER(C.c_int64)), ("num_algorithm", C.c_int32),
                ("PBFICnum", C.POINTER(C.c_int64)), ("num_PBFICnum", C.c_int32)]


class _BilateralInfo(C.Structure):
    _fields_ = [("sigmaS", C.c_double * 3), ("sigmaR", C.c_double * 3), ("process", C.c_int32 * 3), ("algorithm", C.c_int32 * 3),
                ("PBFICnum", C.c_uint32 * 3), ("radius", C.c_uint32 * 3), ("samples", C.c_uint32 * 3), ("step", C.c_uint32 * 3),
                ("exact_lut", C.c_int32 * 3)]


class _MinMaxArgs(C.Structure):
    _fields_ = [("has_minthr", C.c_int32), ("minthr", C.c_double), ("has_maxthr", C.c_int32), ("maxthr", C.c_double),
                ("planes", C.POINTER(C.c_int64)), ("num_planes", C.c_int32)]


class _MinMaxProps(C.Structure):
    _fields_ = [("count", C.c_int32), ("plane", C.c_int32 * 3), ("is_float", C.c_int32), ("has_diff", C.c_int32),
                ("imin", C.c_int64 * 3), ("imax", C.c_int64 * 3), ("fmin", C.c_double * 3), ("fmax", C.c_double * 3), ("diff", C.c_double * 3)]


class _AverageArgs(C.Structure):
    _fields_ = [("exclude", C.POINTER(C.c_int64)), ("num_exclude", C.c_int32), ("planes", C.POINTER(C.c_int64)), ("num_planes", C.c_int32)]


class _LimiterArgs(C.Structure):
    _fields_ = [("min", C.POINTER(C.c_double)), ("num_min", C.c_int32), ("max", C.POINTER(C.c_double)), ("num_max", C.c_int32),
                ("planes", C.POINTER(C.c_int64)), ("num_planes", C.c_int32), ("has_tv_range", C.c_int32), ("tv_range", C.c_int32),
                ("has_mask", C.c_int32), ("mask", C.c_int32)]


class _LimitFilterArgs(C.Structure):
    _fields_ = [("dark_thr", C.POINTER(C.c_double)), ("num_dark_thr", C.c_int32), ("bright_thr", C.POINTER(C.c_double)), ("num_bright_thr", C.c_int32),
                ("elast", C.POINTER(C.c_double)), ("num_elast", C.c_int32), ("planes", C.POINTER(C.c_int64)), ("num_planes", C.c_int32),
                ("color_range", C.c_int32)]


class _AdaptiveBinarizeArgs(C.Structure):
    _fields_ = [("has_c", C.c_int32), ("c", C.c_int64)]


class _AverageProps(C.Structure):
    _fields_ = [("count", C.c_int32), ("plane", C.c_int32 * 3), ("has_diff", C.c_int32), ("avg", C.c_double * 3), ("diff", C.c_double * 3)]


# every symbol include/vszip_cuda.h declares: (restype, argtypes)
_P = C.c_void_p
ABI = {
    "vszip_cuda_init": (C.c_int, [C.POINTER(C.c_int32), C.c_int32]),
    "vszip_cuda_shutdown": (None, []),
    "vszip_cuda_device_count": (C.c_int, []),
    "vszip_cuda_last_error": (C.c_char_p, []),
    "vszip_cuda_abi_version": (C.c_int, []),
    "vszip_cuda_kernel_launches": (C.c_uint64, []),
    "vszip_cuda_stream_sync": (C.c_int, [C.c_int32, _P]),
    "vszip_cuda_host_forget": (None, [_P]),
    "vszip_cuda_host_registered_bytes": (C.c_size_t, []),
    "vszip_cuda_host_register_limit": (C.c_size_t, [C.c_size_t]),
    "vszip_cuda_reserve": (C.c_int32, [C.POINTER(_VideoInfo), C.c_int32]),
    "vszip_boxblur_create": (_P, [C.POINTER(_VideoInfo), C.POINTER(_BoxBlurArgs)]),
    "vszip_boxblur_get_frame": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame), C.POINTER(_Frame)]),
    "vszip_boxblur_device": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "vszip_bilateral_create": (_P, [C.POINTER(_VideoInfo), C.POINTER(_VideoInfo), C.POINTER(_BilateralArgs)]),
    "vszip_bilateral_get_frame": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame), C.POINTER(_Frame), C.POINTER(_Frame)]),
    "vszip_bilateral_get_info": (C.c_int, [_P, C.POINTER(_BilateralInfo)]),
    "vszip_bilateral_device": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "vszip_planeminmax_create": (_P, [C.POINTER(_VideoInfo), C.POINTER(_VideoInfo), C.POINTER(_MinMaxArgs)]),
    "vszip_planeminmax_get_frame": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame), C.POINTER(_Frame), C.POINTER(_MinMaxProps)]),
    "vszip_planeminmax_device": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.POINTER(_MinMaxProps), _P]),
    "vszip_planeaverage_create": (_P, [C.POINTER(_VideoInfo), C.POINTER(_VideoInfo), C.POINTER(_AverageArgs)]),
    "vszip_planeaverage_get_frame": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame), C.POINTER(_Frame), C.POINTER(_AverageProps)]),
    "vszip_planeaverage_device": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.POINTER(_AverageProps), _P]),
    "vszip_planestats_device": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.POINTER(_MinMaxProps), C.POINTER(_AverageProps), C.POINTER(C.c_int32), _P]),
    "vszip_filter_free": (None, [_P]),
    "vszip_filter_planes": (C.c_int, [_P, C.POINTER(C.c_int32 * 3)]),
    "vszip_dev_clip_alloc": (_P, [C.POINTER(_VideoInfo), C.c_int32, C.c_int32]),
    "vszip_dev_clip_free": (None, [_P]),
    "vszip_dev_clip_upload": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame)]),
    "vszip_dev_clip_download": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame)]),
    "vszip_dev_clip_fill_noise": (C.c_int, [_P, C.c_uint64, C.c_int32, C.c_int32]),
    "vszip_dev_clip_frame_bytes": (C.c_size_t, [_P]),
    "vszip_dev_clip_plane_ptr": (_P, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_ssize_t)]),
    "vszip_limiter_create": (_P, [C.POINTER(_VideoInfo), C.POINTER(_LimiterArgs)]),
    "vszip_limiter_get_frame": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame), C.POINTER(_Frame)]),
    "vszip_limiter_device": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "vszip_limitfilter_create": (_P, [C.POINTER(_VideoInfo), C.POINTER(_VideoInfo), C.POINTER(_VideoInfo), C.POINTER(_LimitFilterArgs)]),
    "vszip_limitfilter_get_frame": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame), C.POINTER(_Frame), C.POINTER(_Frame), C.POINTER(_Frame)]),
    "vszip_limitfilter_get_info": (C.c_int, [_P, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3)]),
    "vszip_limitfilter_device": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "vszip_adaptivebinarize_create": (_P, [C.POINTER(_VideoInfo), C.POINTER(_VideoInfo), C.POINTER(_AdaptiveBinarizeArgs)]),
    "vszip_adaptivebinarize_get_frame": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame), C.POINTER(_Frame), C.POINTER(_Frame)]),
    "vszip_adaptivebinarize_device": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "vszip_chain_create": (_P, [C.POINTER(_P), C.c_int32]),
    "vszip_chain_free": (None, [_P]),
    "vszip_chain_planes": (C.c_int, [_P, C.POINTER(C.c_int32 * 3)]),
    "vszip_chain_get_frame": (C.c_int, [_P, C.c_int32, C.POINTER(_Frame), C.POINTER(_Frame), C.POINTER(_P)]),
}

_lib = None


def load_library():
    """dlopen the C-ABI library and bind every declared symbol.  Fails loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise Error(f"{LIB_PATH} is missing: build the CUDA extension with `make` (or __graft_entry__.build()); "
                    "vapoursynth_zip_b200 has no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError = ABI mismatch between header and library
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _last_error() -> str:
    return (load_library().vszip_cuda_last_error() or b"").decode("utf-8", "replace")


# --------------------------------------------------------------------------- formats
@dataclass(frozen=True)
class VideoFormat:
    name: str
    color_family: int
    sample_type: int
    bits_per_sample: int
    bytes_per_sample: int
    subsampling_w: int
    subsampling_h: int
    num_planes: int

    @property
    def dtype(self):
        if self.sample_type == FLOAT:
            return {2: np.float16, 4: np.float32}[self.bytes_per_sample]
        return {1: np.uint8, 2: np.uint16, 4: np.uint32}[self.bytes_per_sample]

    def plane_shape(self, width, height, plane):
        sw, sh = (self.subsampling_w, self.subsampling_h) if plane else (0, 0)
        return (height >> sh, width >> sw)


def _mk(name, fam, st, bits, ssw=0, ssh=0):
    bps = 1 if bits <= 8 else (2 if bits <= 16 else 4)
    return VideoFormat(name, fam, st, bits, bps, ssw, ssh, 1 if fam == GRAY else 3)


FORMATS = {f.name: f for f in [
    _mk("GRAY8", GRAY, INTEGER, 8), _mk("GRAY9", GRAY, INTEGER, 9), _mk("GRAY10", GRAY, INTEGER, 10), _mk("GRAY11", GRAY, INTEGER, 11),
    _mk("GRAY12", GRAY, INTEGER, 12), _mk("GRAY14", GRAY, INTEGER, 14), _mk("GRAY16", GRAY, INTEGER, 16),
    _mk("GRAY32", GRAY, INTEGER, 32), _mk("GRAYH", GRAY, FLOAT, 16), _mk("GRAYS", GRAY, FLOAT, 32),
    _mk("YUV420P8", YUV, INTEGER, 8, 1, 1), _mk("YUV420P10", YUV, INTEGER, 10, 1, 1), _mk("YUV420P16", YUV, INTEGER, 16, 1, 1),
    _mk("YUV420PH", YUV, FLOAT, 16, 1, 1), _mk("YUV420PS", YUV, FLOAT, 32, 1, 1),
    _mk("YUV422P16", YUV, INTEGER, 16, 1, 0),
    _mk("YUV444P8", YUV, INTEGER, 8), _mk("YUV444P10", YUV, INTEGER, 10), _mk("YUV444P16", YUV, INTEGER, 16),
    _mk("YUV444PH", YUV, FLOAT, 16), _mk("YUV444PS", YUV, FLOAT, 32),
    _mk("RGB24", RGB, INTEGER, 8), _mk("RGB30", RGB, INTEGER, 10), _mk("RGB48", RGB, INTEGER, 16),
    _mk("RGBH", RGB, FLOAT, 16), _mk("RGBS", RGB, FLOAT, 32),
]}
globals().update(FORMATS)  # vz.GRAY16, vz.YUV420P16, ... like the vs.* presets


def _fmt(f) -> VideoFormat:
    return f if isinstance(f, VideoFormat) else FORMATS[f]


def _vi(fmt: VideoFormat, width, height, num_frames) -> _VideoInfo:
    return _VideoInfo(width, height, num_frames, fmt.color_family, fmt.sample_type, fmt.bits_per_sample, fmt.bytes_per_sample,
                      fmt.subsampling_w, fmt.subsampling_h, fmt.num_planes)


def _rows(p):
    """Planes are passed with their own row stride; only the samples of a row must be contiguous."""
    return p if (p.ndim == 2 and p.strides[1] == p.itemsize and p.strides[0] > 0) else np.ascontiguousarray(p)


def _cframe(planes) -> _Frame:
    fr = _Frame()
    fr._keep = list(planes)  # the struct only holds raw pointers: keep the arrays alive with it
    for i, p in enumerate(planes):
        if p is None:
            continue
        assert p.ndim == 2 and p.strides[1] == p.itemsize, "plane rows must be contiguous"
        fr.data[i] = p.ctypes.data
        fr.stride[i] = p.strides[0]
    return fr


# --------------------------------------------------------------------------- clip model
class VideoFrame:
    def __init__(self, fmt: VideoFormat, width: int, height: int, planes, props=None):
        self.format, self.width, self.height = fmt, width, height
        self.planes = list(planes)
        self.props = dict(props or {})

    def __getitem__(self, plane):
        return self.planes[plane]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class VideoNode:
    """A clip: format + geometry + a frame getter.  `clip.vszip.X(...)` mirrors the plugin namespace."""

    def __init__(self, fmt: VideoFormat, width: int, height: int, num_frames: int, getter):
        self.format, self.width, self.height, self.num_frames = fmt, width, height, num_frames
        self._getter = getter

    def get_frame(self, n: int) -> VideoFrame:
        if not 0 <= n < self.num_frames:
            raise Error(f"frame {n} out of range")
        chain = _fusable_chain(self) if core.fuse_chains else None
        if chain is not None:
            return _fused_get(chain, n)
        return self._getter(n)

    @property
    def vszip(self):
        return _Namespace(self)

    def _info(self) -> _VideoInfo:
        return _vi(self.format, self.width, self.height, self.num_frames)


def _i64(values):
    vals = [int(v) for v in values]
    return (C.c_int64 * max(len(vals), 1))(*vals), len(vals)


def _f64(values):
    vals = [float(v) for v in values]
    return (C.c_double * max(len(vals), 1))(*vals), len(vals)


def _as_list(v):
    if v is None:
        return None
    return list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v]


_live_filters = weakref.WeakSet()


class _Filter:
    """Owns one vszip_filter handle."""

    def __init__(self, handle, node: VideoNode | None = None):
        if not handle:
            raise Error(_last_error())
        self.handle = handle
        mask = (C.c_int32 * 3)()
        load_library().vszip_filter_planes(handle, C.byref(mask))
        self.process = [bool(m) for m in mask]
        _live_filters.add(self)

    def close(self):
        """vszip_filter_free (xxxFree in the reference's glue); also called by core.shutdown() for every instance still alive."""
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.vszip_filter_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check(rc):
    if rc != 0:
        raise Error(_last_error())


def _planes_arg(planes):
    if planes is None:
        return None, -1
    return _i64(_as_list(planes))


# ---- BoxBlur
class BoxBlurFilter(_Filter):
    def __init__(self, vi: _VideoInfo, planes=None, hradius=None, hpasses=None, vradius=None, vpasses=None):
        pl, npl = _planes_arg(planes)
        a = _BoxBlurArgs(pl, npl, hradius is not None, int(hradius or 0), hpasses is not None, int(hpasses or 0),
                         vradius is not None, int(vradius or 0), vpasses is not None, int(vpasses or 0))
        super().__init__(load_library().vszip_boxblur_create(C.byref(vi), C.byref(a)))

    def run_device(self, src: "DeviceClip", dst: "DeviceClip", first=0, count=None, stream=None):
        count = src.num_frames - first if count is None else count
        _check(load_library().vszip_boxblur_device(self.handle, src.handle, dst.handle, first, count, stream))


class BilateralFilter(_Filter):
    def __init__(self, vi: _VideoInfo, ref_vi: _VideoInfo | None = None, sigmaS=None, sigmaR=None, planes=None, algorithm=None, PBFICnum=None):
        s, ns = _f64(_as_list(sigmaS) or [])
        r, nr = _f64(_as_list(sigmaR) or [])
        al, na = _i64(_as_list(algorithm) or [])
        pb, npb = _i64(_as_list(PBFICnum) or [])
        pl, npl = _planes_arg(planes)
        a = _BilateralArgs(s, ns, r, nr, pl, npl, al, na, pb, npb)
        super().__init__(load_library().vszip_bilateral_create(C.byref(vi), C.byref(ref_vi) if ref_vi is not None else None, C.byref(a)))

    def info(self):
        out = _BilateralInfo()
        _check(load_library().vszip_bilateral_get_info(self.handle, C.byref(out)))
        return out

    def run_device(self, src, dst, ref=None, first=0, count=None, stream=None):
        count = src.num_frames - first if count is None else count
        _check(load_library().vszip_bilateral_device(self.handle, src.handle, ref.handle if ref else None, dst.handle, first, count, stream))


def _props_from(prop, keys, count, getter):
    """Append semantics (src/vapoursynth/planeminmax.zig:48-72): scalar for one plane, list for several."""
    out = {}
    for key in keys:
        vals = [getter(key, i) for i in range(count)]
        out[prop + key] = vals[0] if len(vals) == 1 else vals
    return out


class PlaneMinMaxFilter(_Filter):
    def __init__(self, vi: _VideoInfo, clipb_vi: _VideoInfo | None = None, minthr=None, maxthr=None, planes=None):
        pl, npl = _planes_arg(planes)
        a = _MinMaxArgs(minthr is not None, float(minthr or 0), maxthr is not None, float(maxthr or 0), pl, npl)
        super().__init__(load_library().vszip_planeminmax_create(C.byref(vi), C.byref(clipb_vi) if clipb_vi is not None else None, C.byref(a)))

    @staticmethod
    def to_props(p: _MinMaxProps, prop="psm"):
        if p.count == 0:
            return {}
        def get(key, i):
            if key == "Diff":
                return p.diff[i]
            if p.is_float:
                return p.fmin[i] if key == "Min" else p.fmax[i]
            return int(p.imin[i]) if key == "Min" else int(p.imax[i])
        return _props_from(prop, ("Min", "Max") + (("Diff",) if p.has_diff else ()), p.count, get)

    def run_device(self, clipa, clipb=None, first=0, count=None, stream=None, prop="psm"):
        count = clipa.num_frames - first if count is None else count
        out = (_MinMaxProps * max(count, 1))()
        _check(load_library().vszip_planeminmax_device(self.handle, clipa.handle, clipb.handle if clipb else None, first, count, out, stream))
        return [self.to_props(out[i], prop) for i in range(count)]


class PlaneAverageFilter(_Filter):
    def __init__(self, vi: _VideoInfo, clipb_vi: _VideoInfo | None = None, exclude=None, planes=None):
        pl, npl = _planes_arg(planes)
        if exclude is None:
            ex, nex = None, -1
        else:
            ex, nex = _i64(_as_list(exclude))
        a = _AverageArgs(ex, nex, pl, npl)
        super().__init__(load_library().vszip_planeaverage_create(C.byref(vi), C.byref(clipb_vi) if clipb_vi is not None else None, C.byref(a)))

    @staticmethod
    def to_props(p: _AverageProps, prop="psm"):
        if p.count == 0:
            return {}
        out = _props_from(prop, ("Avg",), p.count, lambda k, i: p.avg[i])
        if p.has_diff:
            out.update(_props_from(prop, ("Diff",), p.count, lambda k, i: p.diff[i]))
        return out

    def run_device(self, clipa, clipb=None, first=0, count=None, stream=None, prop="psm"):
        count = clipa.num_frames - first if count is None else count
        out = (_AverageProps * max(count, 1))()
        _check(load_library().vszip_planeaverage_device(self.handle, clipa.handle, clipb.handle if clipb else None, first, count, out, stream))
        return [self.to_props(out[i], prop) for i in range(count)]


def plane_stats_device(minmax: "PlaneMinMaxFilter", average: "PlaneAverageFilter", clipa, first=0, count=None, stream=None, prop="psm", fetch=True):
    """PlaneMinMax + PlaneAverage over the same device-resident frames (vszip_planestats_device): one read of each plane when the pair
    is eligible.  Returns (props per frame - both filters' props merged, like `clip.vszip.PlaneMinMax().vszip.PlaneAverage()` -, fused)."""
    count = clipa.num_frames - first if count is None else count
    mm = (_MinMaxProps * max(count, 1))() if fetch else None
    av = (_AverageProps * max(count, 1))() if fetch else None
    fused = C.c_int32(0)
    _check(load_library().vszip_planestats_device(minmax.handle, average.handle, clipa.handle, first, count, mm, av, C.byref(fused), stream))
    if not fetch:
        return None, bool(fused.value)
    out = []
    for i in range(count):
        p = PlaneMinMaxFilter.to_props(mm[i], prop)
        p.update(PlaneAverageFilter.to_props(av[i], prop))
        out.append(p)
    return out, bool(fused.value)


class LimiterFilter(_Filter):
    def __init__(self, vi: _VideoInfo, min=None, max=None, tv_range=None, mask=None, planes=None):
        pl, npl = _planes_arg(planes)
        mn, nmn = (None, -1) if min is None else _f64(_as_list(min))
        mx, nmx = (None, -1) if max is None else _f64(_as_list(max))
        a = _LimiterArgs(mn, nmn, mx, nmx, pl, npl, tv_range is not None, int(bool(tv_range)), mask is not None, int(bool(mask)))
        super().__init__(load_library().vszip_limiter_create(C.byref(vi), C.byref(a)))

    def run_device(self, src: "DeviceClip", dst: "DeviceClip", first=0, count=None, stream=None):
        count = src.num_frames - first if count is None else count
        _check(load_library().vszip_limiter_device(self.handle, src.handle, dst.handle, first, count, stream))


class LimitFilterFilter(_Filter):
    """vszip.LimitFilter (src/vapoursynth/limit_filter.zig:93-124).  color_range: the flt clip's _ColorRange (0 full, 1 limited, None absent)."""

    def __init__(self, flt_vi: _VideoInfo, src_vi: _VideoInfo, ref_vi: _VideoInfo | None = None, dark_thr=None, bright_thr=None, elast=None,
                 planes=None, color_range=None):
        dk, ndk = _f64(_as_list(dark_thr) or [])
        br, nbr = _f64(_as_list(bright_thr) or [])
        el, nel = _f64(_as_list(elast) or [])
        pl, npl = _planes_arg(planes)
        a = _LimitFilterArgs(dk, ndk, br, nbr, el, nel, pl, npl, -1 if color_range is None else int(color_range))
        super().__init__(load_library().vszip_limitfilter_create(C.byref(flt_vi), C.byref(src_vi) if src_vi is not None else None,
                                                                 C.byref(ref_vi) if ref_vi is not None else None, C.byref(a)))

    def info(self):
        d, b, e = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
        _check(load_library().vszip_limitfilter_get_info(self.handle, C.byref(d), C.byref(b), C.byref(e)))
        return {"dark_thr": list(d), "bright_thr": list(b), "elast": list(e)}

    def run_device(self, flt, src, dst, ref=None, first=0, count=None, stream=None):
        count = flt.num_frames - first if count is None else count
        _check(load_library().vszip_limitfilter_device(self.handle, flt.handle, src.handle, ref.handle if ref else None, dst.handle, first, count, stream))


class AdaptiveBinarizeFilter(_Filter):
    """vszip.AdaptiveBinarize (src/vapoursynth/adaptive_binarize.zig:79-116)."""

    def __init__(self, vi: _VideoInfo, clip2_vi: _VideoInfo, c=None):
        a = _AdaptiveBinarizeArgs(c is not None, int(c or 0))
        super().__init__(load_library().vszip_adaptivebinarize_create(C.byref(vi), C.byref(clip2_vi) if clip2_vi is not None else None, C.byref(a)))

    def run_device(self, clip, clip2, dst, first=0, count=None, stream=None):
        count = clip.num_frames - first if count is None else count
        _check(load_library().vszip_adaptivebinarize_device(self.handle, clip.handle, clip2.handle, dst.handle, first, count, stream))


class _Namespace:
    """`clip.vszip` / `core.vszip`: the plugin functions with the reference's argument names."""

    def __init__(self, clip: VideoNode | None = None):
        self._clip = clip

    def _c(self, clip):
        clip = clip if clip is not None else self._clip
        if clip is None:
            raise Error("clip argument is required")
        return clip

    def BoxBlur(self, clip=None, planes=None, hradius=None, hpasses=None, vradius=None, vpasses=None) -> VideoNode:
        clip = self._c(clip)
        flt = BoxBlurFilter(clip._info(), planes, hradius, hpasses, vradius, vpasses)
        return _pixel_node(clip, None, flt, lambda n, s, r, d: load_library().vszip_boxblur_get_frame(flt.handle, n, C.byref(s), C.byref(d)))

    def Limiter(self, clip=None, min=None, max=None, tv_range=None, mask=None, planes=None) -> VideoNode:
        clip = self._c(clip)
        flt = LimiterFilter(clip._info(), min, max, tv_range, mask, planes)
        return _pixel_node(clip, None, flt, lambda n, s, r, d: load_library().vszip_limiter_get_frame(flt.handle, n, C.byref(s), C.byref(d)))

    def _bind(self, names, pos, kw):
        """VapourSynth binding rule: `clip.vszip.F(a, b)` puts the bound clip first and shifts the positional arguments."""
        if self._clip is not None:
            pos = (self._clip,) + tuple(pos)
        if len(pos) > len(names):
            raise Error("too many positional arguments")
        args = dict(zip(names, pos))
        for k, v in kw.items():
            if k not in names:
                raise Error(f"unknown argument {k}")
            if k in args:
                raise Error(f"argument {k} given twice")
            args[k] = v
        return [args.get(k) for k in names]

    def LimitFilter(self, *pos, **kw) -> VideoNode:
        flt, src, ref, dark_thr, bright_thr, elast, planes = self._bind(("flt", "src", "ref", "dark_thr", "bright_thr", "elast", "planes"), pos, kw)
        if flt is None or src is None:
            raise Error("LimitFilter: flt and src are required")
        f = LimitFilterFilter(flt._info(), src._info(), ref._info() if ref is not None else None, dark_thr, bright_thr, elast, planes, None)
        # hz.scaleValue -> hz.getColorRange reads frame 0 of flt (src/helper.zig:259-276), only when the depth is not 8 bits
        if flt.format.bits_per_sample != 8 and flt.num_frames > 0:
            cr = flt.get_frame(0).props.get("_ColorRange")
            if cr is not None:
                f = LimitFilterFilter(flt._info(), src._info(), ref._info() if ref is not None else None, dark_thr, bright_thr, elast, planes, cr)
        return _multi_node(flt, [src] + ([ref] if ref is not None else []), f, lambda n, a, others, d: load_library().vszip_limitfilter_get_frame(
            f.handle, n, C.byref(a), C.byref(others[0]), C.byref(others[1]) if len(others) > 1 else None, C.byref(d)),
            through=flt if ref is None else None, needs_source=src)

    def AdaptiveBinarize(self, *pos, **kw) -> VideoNode:
        clip, clip2, c = self._bind(("clip", "clip2", "c"), pos, kw)
        if clip is None or clip2 is None:
            raise Error("AdaptiveBinarize: clip and clip2 are required")
        f = AdaptiveBinarizeFilter(clip._info(), clip2._info(), c)
        return _multi_node(clip, [clip2], f, lambda n, a, others, d: load_library().vszip_adaptivebinarize_get_frame(
            f.handle, n, C.byref(a), C.byref(others[0]), C.byref(d)), {"_ColorRange": 0},  # dst_prop.setColorRange(.FULL)
            through=clip2, needs_source=clip)

    def Bilateral(self, clip=None, ref=None, sigmaS=None, sigmaR=None, planes=None, algorithm=None, PBFICnum=None) -> VideoNode:
        clip = self._c(clip)
        flt = BilateralFilter(clip._info(), ref._info() if ref is not None else None, sigmaS, sigmaR, planes, algorithm, PBFICnum)
        return _pixel_node(clip, ref, flt, lambda n, s, r, d: load_library().vszip_bilateral_get_frame(
            flt.handle, n, C.byref(s), C.byref(r) if r is not None else None, C.byref(d)))

    def PlaneMinMax(self, clipa=None, minthr=None, maxthr=None, clipb=None, planes=None, prop=None) -> VideoNode:
        clipa = self._c(clipa)
        flt = PlaneMinMaxFilter(clipa._info(), clipb._info() if clipb is not None else None, minthr, maxthr, planes)
        prefix = "psm" if prop is None else str(prop)

        def run(n, a, b):
            out = _MinMaxProps()
            _check(load_library().vszip_planeminmax_get_frame(flt.handle, n, C.byref(a), C.byref(b) if b is not None else None, C.byref(out)))
            return PlaneMinMaxFilter.to_props(out, prefix)
        return _props_node(clipa, clipb, flt, run, [prefix + k for k in ("Diff", "Max", "Min")], lambda o: PlaneMinMaxFilter.to_props(o, prefix))

    def PlaneAverage(self, clipa=None, exclude=None, clipb=None, planes=None, prop=None) -> VideoNode:
        clipa = self._c(clipa)
        flt = PlaneAverageFilter(clipa._info(), clipb._info() if clipb is not None else None, exclude, planes)
        prefix = "psm" if prop is None else str(prop)

        def run(n, a, b):
            out = _AverageProps()
            _check(load_library().vszip_planeaverage_get_frame(flt.handle, n, C.byref(a), C.byref(b) if b is not None else None, C.byref(out)))
            return PlaneAverageFilter.to_props(out, prefix)
        return _props_node(clipa, clipb, flt, run, [prefix + k for k in ("Diff", "Avg")], lambda o: PlaneAverageFilter.to_props(o, prefix))


def _second_frame(second: VideoNode | None, n: int):
    if second is None:
        return None
    return second.get_frame(min(n, second.num_frames - 1))  # rpFrameReuseLastOnly for a shorter-or-equal second clip


def _pixel_node(clip: VideoNode, ref: VideoNode | None, flt: _Filter, call) -> VideoNode:
    fmt = clip.format

    def get(n):
        core._ensure_init()
        src = clip.get_frame(n)
        rf = _second_frame(ref, n)
        # newVideoFrame2(planes): processed planes are fresh, the others are shared with the source
        out_planes = [np.empty_like(p, order="C") if flt.process[i] else p for i, p in enumerate(src.planes)]
        s = _cframe([_rows(p) if flt.process[i] else None for i, p in enumerate(src.planes)])
        keep = [_rows(p) for p in rf.planes] if rf is not None else None
        r = _cframe(keep) if rf is not None else None
        d = _cframe([p if flt.process[i] else None for i, p in enumerate(out_planes)])
        _check(call(n, s, r, d))
        return VideoFrame(fmt, clip.width, clip.height, out_planes, src.props)

    node = VideoNode(fmt, clip.width, clip.height, clip.num_frames, get)
    node.filter = flt
    node._chain_info = ("pixel", clip, ref, None)
    return node


def _multi_node(first: VideoNode, others, flt: _Filter, call, set_props=None, through=None, needs_source=None) -> VideoNode:
    """A pixel filter with several input clips (LimitFilter, AdaptiveBinarize): dst is allocated from `first`."""
    fmt = first.format

    def get(n):
        core._ensure_init()
        src = first.get_frame(n)
        ofr = [_second_frame(o, n) for o in others]
        out_planes = [np.empty_like(p, order="C") if flt.process[i] else p for i, p in enumerate(src.planes)]
        keep = [[_rows(p) if flt.process[i] else None for i, p in enumerate(fr.planes)] for fr in [src] + ofr]
        d = _cframe([p if flt.process[i] else None for i, p in enumerate(out_planes)])
        _check(call(n, _cframe(keep[0]), [_cframe(k) for k in keep[1:]], d))
        props = dict(src.props)
        props.update(set_props or {})
        return VideoFrame(fmt, first.width, first.height, out_planes, props)

    node = VideoNode(fmt, first.width, first.height, first.num_frames, get)
    node.filter = flt
    # "diamond" element of a fused chain: `through` is the input that continues the chain, `needs_source` the clip that must turn
    # out to be the chain's source (checked in _fusable_chain); anything else (a ref clip, ...) is never fused
    node._chain_info = ("pixel", through, None, None) if through is not None else None
    node._needs_source = needs_source
    node._set_props = set_props
    node._props_from_source = needs_source is not None and first is needs_source   # dst = first.newVideoFrame(): props of `first` only
    return node


def _props_node(clipa: VideoNode, clipb: VideoNode | None, flt: _Filter, run, drop_keys, run_fused=None) -> VideoNode:
    def get(n):
        core._ensure_init()
        src = clipa.get_frame(n)
        bf = _second_frame(clipb, n)
        a = _cframe([_rows(p) if flt.process[i] else None for i, p in enumerate(src.planes)])
        keep = [_rows(p) for p in bf.planes] if bf is not None else None
        b = _cframe(keep) if bf is not None else None
        props = dict(src.props)                      # copyFrame: pixels shared, props copied
        for k in drop_keys:
            props.pop(k, None)
        props.update(run(n, a, b))
        return VideoFrame(src.format, src.width, src.height, src.planes, props)

    node = VideoNode(clipa.format, clipa.width, clipa.height, clipa.num_frames, get)
    node.filter = flt
    node._chain_info = ("props", clipa, clipb, (run_fused, drop_keys))
    return node


# ---- fused evaluation of linear chains of vszip nodes (vszip_chain_*): one upload, one download
class _Chain:
    def __init__(self, nodes):
        self.nodes = nodes  # evaluation order
        handles = (_P * len(nodes))(*[nd.filter.handle for nd in nodes])
        self.handle = load_library().vszip_chain_create(handles, len(nodes))
        if not self.handle:
            raise Error(_last_error())
        w = (C.c_int32 * 3)()
        _check(load_library().vszip_chain_planes(self.handle, C.byref(w)))
        self.written = [bool(v) for v in w]

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.vszip_chain_free(self.handle)
        except Exception:
            pass
        self.handle = None


def _fusable_chain(node: "VideoNode"):
    """The chain ending in `node` if it and at least one predecessor are single-input vszip nodes."""
    cached = getattr(node, "_chain", False)
    if cached is not False:
        return cached
    nodes, cur = [], node
    while getattr(cur, "_chain_info", None) is not None and cur._chain_info[2] is None and len(nodes) < 16:
        nodes.append(cur)
        cur = cur._chain_info[1]
    # diamond elements (LimitFilter's src, AdaptiveBinarize's clip) must read exactly the chain's source
    ok = all(getattr(nd, "_needs_source", None) is None or nd._needs_source is cur for nd in nodes)
    node._chain = (_Chain(nodes[::-1]), cur) if len(nodes) >= 2 and ok else None
    return node._chain


def _fused_get(chain_and_source, n: int) -> "VideoFrame":
    chain, source = chain_and_source
    core._ensure_init()
    src = source.get_frame(n)
    out_planes = [np.empty_like(p, order="C") if chain.written[i] else p for i, p in enumerate(src.planes)]
    s = _cframe([_rows(p) for p in src.planes])
    d = _cframe([p if chain.written[i] else None for i, p in enumerate(out_planes)])
    outs, ptrs = [], (_P * len(chain.nodes))()
    for i, nd in enumerate(chain.nodes):
        kind = nd._chain_info[0]
        o = None
        if kind == "props":
            o = _MinMaxProps() if isinstance(nd.filter, PlaneMinMaxFilter) else _AverageProps()
            ptrs[i] = C.cast(C.pointer(o), _P)
        outs.append(o)
    _check(load_library().vszip_chain_get_frame(chain.handle, n, C.byref(s), C.byref(d), ptrs))
    props = dict(src.props)
    for nd, o in zip(chain.nodes, outs):        # evaluation order, exactly what the unfused nodes would do to the props
        if o is not None:
            to_props, drop_keys = nd._chain_info[3]
            for k in drop_keys:
                props.pop(k, None)
            props.update(to_props(o))
        if getattr(nd, "_props_from_source", False):
            props = dict(src.props)             # e.g. AdaptiveBinarize(src, src.BoxBlur().PlaneMinMax()): clip2's props are not inherited
        props.update(getattr(nd, "_set_props", None) or {})
    return VideoFrame(source.format, source.width, source.height, out_planes, props)


# --------------------------------------------------------------------------- device-resident clips
class DeviceClip:
    """N frames resident in one GPU's HBM (vszip_dev_clip)."""

    def __init__(self, fmt, width, height, num_frames, device=0):
        core._ensure_init()
        self.format, self.width, self.height, self.num_frames, self.device = _fmt(fmt), width, height, num_frames, device
        vi = _vi(self.format, width, height, num_frames)
        self.handle = load_library().vszip_dev_clip_alloc(C.byref(vi), num_frames, device)
        if not self.handle:
            raise Error(_last_error())

    def info(self) -> _VideoInfo:
        return _vi(self.format, self.width, self.height, self.num_frames)

    @property
    def frame_bytes(self) -> int:
        return int(load_library().vszip_dev_clip_frame_bytes(self.handle))

    def fill_noise(self, seed=1234, first_frame_no=0, frame_no_stride=1):
        _check(load_library().vszip_dev_clip_fill_noise(self.handle, seed, first_frame_no, frame_no_stride))

    def upload(self, frame: int, planes):
        keep = [np.ascontiguousarray(p) for p in planes]
        _check(load_library().vszip_dev_clip_upload(self.handle, frame, C.byref(_cframe(keep))))

    def download(self, frame: int):
        planes = [np.empty(self.format.plane_shape(self.width, self.height, p), self.format.dtype) for p in range(self.format.num_planes)]
        _check(load_library().vszip_dev_clip_download(self.handle, frame, C.byref(_cframe(planes))))
        return planes

    def plane_ptr(self, frame, plane):
        pitch = C.c_ssize_t()
        return load_library().vszip_dev_clip_plane_ptr(self.handle, frame, plane, C.byref(pitch)), pitch.value

    def free(self):
        if self.handle and _lib is not None:
            _lib.vszip_dev_clip_free(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# --------------------------------------------------------------------------- core
class _Core:
    def __init__(self):
        self._ready = False
        # evaluate linear chains of single-input vszip nodes with one upload and one download (vszip_chain_*)
        self.fuse_chains = os.environ.get("VSZIP_FUSE_CHAINS", "1") != "0"

    def init(self, devices=None) -> int:
        """vszip_cuda_init: `devices` = list of CUDA ordinals (default: all visible, or $VSZIP_CUDA_DEVICES)."""
        lib = load_library()
        if devices is None and os.environ.get("VSZIP_CUDA_DEVICES"):
            devices = [int(x) for x in os.environ["VSZIP_CUDA_DEVICES"].split(",")]
        if devices:
            arr = (C.c_int32 * len(devices))(*devices)
            n = lib.vszip_cuda_init(arr, len(devices))
        else:
            n = lib.vszip_cuda_init(None, 0)
        if n <= 0:
            raise Error(_last_error())
        self._ready = True
        return n

    def reserve(self, info, buffers=2):
        """vszip_cuda_reserve: allocate every request slot's staging buffers for frames of this format now (first-call latency)."""
        self._ensure_init()
        _check(load_library().vszip_cuda_reserve(C.byref(info), buffers))

    def shutdown(self):
        """vszip_cuda_shutdown: frees the per-GPU slots, streams and the host pin cache (filters and clips must be gone)."""
        import gc
        gc.collect()
        for f in list(_live_filters):     # instances the interpreter has not finalised yet: their device tables go first
            f.close()
        load_library().vszip_cuda_shutdown()
        self._ready = False

    def _ensure_init(self):
        if not self._ready:
            self.init()

    @property
    def num_devices(self):
        return load_library().vszip_cuda_device_count()

    @property
    def kernel_launches(self) -> int:
        return int(load_library().vszip_cuda_kernel_launches())

    def sync(self, device=0, stream=None):
        _check(load_library().vszip_cuda_stream_sync(device, stream))

    @property
    def vszip(self):
        return _Namespace(None)

    # ---- sources (stand-ins for std.BlankClip / a decoded clip)
    def clip_from_frames(self, fmt, frames, props=None) -> VideoNode:
        """frames: list (one entry per frame) of lists of 2-D numpy planes."""
        fmt = _fmt(fmt)
        h, w = frames[0][0].shape
        for fr in frames:
            assert len(fr) == fmt.num_planes
            for i, p in enumerate(fr):
                assert p.dtype == fmt.dtype and p.shape == fmt.plane_shape(w, h, i), (p.dtype, p.shape, fmt.name)
        return VideoNode(fmt, w, h, len(frames), lambda n: VideoFrame(fmt, w, h, frames[n], props))

    def BlankClip(self, fmt, width, height, length=1, color=None) -> VideoNode:
        fmt = _fmt(fmt)
        color = _as_list(color) or [0] * fmt.num_planes
        color = color + [color[-1]] * (fmt.num_planes - len(color))
        planes = [np.full(fmt.plane_shape(width, height, i), color[i], dtype=fmt.dtype) for i in range(fmt.num_planes)]
        return VideoNode(fmt, width, height, length, lambda n: VideoFrame(fmt, width, height, planes))


core = _Core()
