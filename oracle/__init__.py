"""ctypes binding of the CPU oracle (oracle/vszip_oracle.cpp, oracle/fixture.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package never imports this module.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "libvszip_oracle.so"

U8, U16, F16, F32, U32 = 0, 1, 2, 3, 4   # U32: Limiter only
_NP = {U8: np.uint8, U16: np.uint16, F16: np.float16, F32: np.float32, U32: np.uint32}


def sample_type_of(arr: np.ndarray) -> int:
    for k, v in _NP.items():
        if arr.dtype == v:
            return k
    raise TypeError(f"unsupported dtype {arr.dtype}")


def build(force: bool = False) -> Path:
    srcs = [_DIR / "vszip_oracle.cpp", _DIR / "fixture.cpp"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


class _BilateralParams(C.Structure):
    _fields_ = [
        ("sigmaS", C.c_double * 3), ("sigmaR", C.c_double * 3),
        ("process", C.c_int * 3), ("algorithm", C.c_int * 3),
        ("pbfic_num", C.c_uint * 3), ("radius", C.c_uint * 3), ("samples", C.c_uint * 3), ("step", C.c_uint * 3),
        ("peak", C.c_float), ("hist_len", C.c_int),
    ]


class _MinMaxOut(C.Structure):
    _fields_ = [("imin", C.c_longlong), ("imax", C.c_longlong), ("fmin", C.c_double), ("fmax", C.c_double), ("diff", C.c_double)]


class _AverageOut(C.Structure):
    _fields_ = [("avg", C.c_double), ("diff", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.vso_boxblur_plane.argtypes = [C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t] + [C.c_int] * 6
        _lib.vso_bilateral_plane.argtypes = [C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t,
                                             C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
        _lib.vso_bilateral_pbfic_plane.argtypes = [C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t,
                                                   C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        _lib.vso_recursive_gaussian_params.argtypes = [C.c_double, C.c_void_p]
        _lib.vso_recursive_gaussian_params.restype = None
        _lib.vso_limiter_plane.argtypes = [C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_double, C.c_double]
        _lib.vso_limitfilter_plane.argtypes = [C.c_int] + [C.c_void_p, C.c_ssize_t] * 4 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        _lib.vso_adaptive_binarize_plane.argtypes = [C.c_void_p, C.c_ssize_t] * 3 + [C.c_int] * 3
        _lib.vso_bilateral_luts.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.vso_bilateral_luts.restype = None
        _lib.vso_planeminmax_plane.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int,
                                               C.c_float, C.c_float, C.POINTER(_MinMaxOut)]
        _lib.vso_planeaverage_plane.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int,
                                                C.c_void_p, C.c_int, C.POINTER(_AverageOut)]
        _lib.vso_plane_stats.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 3
        _lib.vso_plane_stats.restype = None
        _lib.vsf_rgb8_to_rgbs.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _lib.vsf_rgbs_to_yuv_plane.argtypes = [C.c_void_p] * 3 + [C.c_size_t, C.c_int, C.c_void_p]
        _lib.vsf_chroma_420.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.vsf_quantise.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        _lib.vsf_f32_to_f16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _chk2d(a: np.ndarray):
    assert a.ndim == 2 and a.strides[1] == a.itemsize, "plane must be 2-D with contiguous rows"


# --------------------------------------------------------------------------- filters

def boxblur_plane(src: np.ndarray, hradius=1, hpasses=1, vradius=1, vpasses=1) -> np.ndarray:
    """vszip.BoxBlur on one plane (CT/RT dispatch as src/vapoursynth/boxblur.zig:188)."""
    _chk2d(src)
    dst = np.empty_like(src, order="C")
    h, w = src.shape
    rc = lib().vso_boxblur_plane(sample_type_of(src), _p(src), src.strides[0], _p(dst), dst.strides[0], w, h,
                                 hradius, hpasses, vradius, vpasses)
    assert rc == 0
    return dst


def bilateral_derive(is_yuv, sample_is_float, bits, ssw, ssh, num_planes, sigmaS=(), sigmaR=(), planes=None,
                     algorithm=(), PBFICnum=()):
    """Per-plane parameter derivation of src/vapoursynth/bilateral.zig:104-199.  Returns (rc, params)."""
    def arr(t, v):
        v = list(v)
        return (t * max(len(v), 1))(*v), len(v)
    s, ns = arr(C.c_double, sigmaS)
    r, nr = arr(C.c_double, sigmaR)
    a, na = arr(C.c_int, algorithm)
    p, np_ = arr(C.c_int, PBFICnum)
    if planes is None:
        pl, npl = (C.c_int * 1)(), -1
    else:
        pl, npl = arr(C.c_int, planes)
    out = _BilateralParams()
    rc = lib().vso_bilateral_derive(int(is_yuv), int(sample_is_float), bits, ssw, ssh, num_planes, s, ns, r, nr,
                                    pl, npl, a, na, p, np_, C.byref(out))
    return rc, out


def bilateral_luts(sigmaS, sigmaR, radius, hist_len):
    gs = np.empty((radius + 1) ** 2, np.float32)
    gr = np.empty(hist_len, np.float32)
    lib().vso_bilateral_luts(sigmaS, sigmaR, radius, hist_len, _p(gs), _p(gr))
    return gs, gr


def bilateral_plane(src: np.ndarray, sigmaS, sigmaR, radius, step, hist_len, ref: np.ndarray | None = None) -> np.ndarray:
    """Algorithm-2 bilateral on one plane with already-derived per-plane parameters."""
    _chk2d(src)
    if ref is None:
        ref = src
    _chk2d(ref)
    assert ref.shape == src.shape and ref.dtype == src.dtype
    dst = np.empty_like(src, order="C")
    h, w = src.shape
    rc = lib().vso_bilateral_plane(sample_type_of(src), _p(src), src.strides[0], _p(ref), ref.strides[0], _p(dst), dst.strides[0],
                                   w, h, float(sigmaS), float(sigmaR), int(radius), int(step), int(hist_len))
    assert rc == 0
    return dst


def bilateral_pbfic_plane(src: np.ndarray, sigmaS, sigmaR, pbfic_num, hist_len, ref: np.ndarray | None = None) -> np.ndarray:
    """Algorithm-1 (PBFIC) bilateral on one plane with already-derived per-plane parameters."""
    _chk2d(src)
    if ref is None:
        ref = src
    _chk2d(ref)
    assert ref.shape == src.shape and ref.dtype == src.dtype
    dst = np.empty_like(src, order="C")
    h, w = src.shape
    rc = lib().vso_bilateral_pbfic_plane(sample_type_of(src), _p(src), src.strides[0], _p(ref), ref.strides[0], _p(dst), dst.strides[0],
                                         w, h, float(sigmaS), float(sigmaR), int(pbfic_num), int(hist_len))
    assert rc == 0
    return dst


def recursive_gaussian_params(sigma) -> np.ndarray:
    out = np.zeros(4, np.float32)
    lib().vso_recursive_gaussian_params(float(sigma), _p(out))
    return out


def limiter_plane(src: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """dst = min(max(lo, src), hi) with the bounds rounded to the sample type."""
    _chk2d(src)
    dst = np.empty_like(src, order="C")
    h, w = src.shape
    rc = lib().vso_limiter_plane(sample_type_of(src), _p(src), src.strides[0], _p(dst), dst.strides[0], w, h, float(lo), float(hi))
    assert rc == 0
    return dst


def limitfilter_plane(flt: np.ndarray, src: np.ndarray, ref: np.ndarray | None, dark_thr: float, bright_thr: float, elast: float) -> np.ndarray:
    """src/filters/limit_filter.zig:3-34; thresholds already scaled to the clip's depth."""
    _chk2d(flt); _chk2d(src)
    ref = src if ref is None else ref
    _chk2d(ref)
    assert flt.shape == src.shape == ref.shape and flt.dtype == src.dtype == ref.dtype
    dst = np.empty_like(flt, order="C")
    h, w = flt.shape
    rc = lib().vso_limitfilter_plane(sample_type_of(flt), _p(flt), flt.strides[0], _p(src), src.strides[0], _p(ref), ref.strides[0],
                                     _p(dst), dst.strides[0], w, h, float(dark_thr), float(bright_thr), float(elast))
    assert rc == 0
    return dst


def adaptive_binarize_plane(a: np.ndarray, b: np.ndarray, c: int) -> np.ndarray:
    """src/vapoursynth/adaptive_binarize.zig:48-60: 255 where b - a >= c."""
    _chk2d(a); _chk2d(b)
    assert a.dtype == np.uint8 and b.dtype == np.uint8 and a.shape == b.shape
    dst = np.empty_like(a, order="C")
    h, w = a.shape
    rc = lib().vso_adaptive_binarize_plane(_p(a), a.strides[0], _p(b), b.strides[0], _p(dst), dst.strides[0], w, h, int(c))
    assert rc == 0
    return dst


def planeminmax_plane(src: np.ndarray, bits: int, minthr=0.0, maxthr=0.0, ref: np.ndarray | None = None) -> dict:
    _chk2d(src)
    out = _MinMaxOut()
    h, w = src.shape
    rp, rs = (None, 0) if ref is None else (_p(ref), ref.strides[0])
    rc = lib().vso_planeminmax_plane(sample_type_of(src), bits, _p(src), src.strides[0], rp, rs, w, h,
                                     float(np.float32(minthr)), float(np.float32(maxthr)), C.byref(out))
    assert rc == 0
    is_float = src.dtype in (np.float16, np.float32)
    res = {"Min": out.fmin if is_float else int(out.imin), "Max": out.fmax if is_float else int(out.imax)}
    if ref is not None:
        res["Diff"] = out.diff
    return res


def planeaverage_plane(src: np.ndarray, bits: int, exclude, ref: np.ndarray | None = None) -> dict:
    _chk2d(src)
    out = _AverageOut()
    h, w = src.shape
    ex = np.asarray(list(exclude), dtype=np.int32)
    rp, rs = (None, 0) if ref is None else (_p(ref), ref.strides[0])
    rc = lib().vso_planeaverage_plane(sample_type_of(src), bits, _p(src), src.strides[0], rp, rs, w, h, _p(ex), len(ex), C.byref(out))
    assert rc == 0
    res = {"Avg": out.avg}
    if ref is not None:
        res["Diff"] = out.diff
    return res


def plane_stats(src: np.ndarray, bits: int) -> dict:
    """std.PlaneStats {avg,min,max} the way the reference's golden store records them."""
    _chk2d(src)
    a, mn, mx = C.c_double(), C.c_double(), C.c_double()
    h, w = src.shape
    lib().vso_plane_stats(sample_type_of(src), bits, _p(src), src.strides[0], w, h, C.byref(a), C.byref(mn), C.byref(mx))
    return {"avg": a.value, "min": mn.value, "max": mx.value}
