"""The reference's filter semantics (plane selection, per-plane parameters, frame props) on top
of the per-plane CPU oracle.  Test infrastructure: mirrors what getFrame does in
src/vapoursynth/{boxblur,bilateral,planeminmax,planeaverage}.zig so parity tests can run the
same call against the oracle and against the CUDA product.

A clip is {"format": name, "planes": [np.ndarray, ...]} (see oracle/fixtures.py).
"""
import numpy as np

import oracle
from oracle.fixtures import FORMATS


def _fmt(clip):
    return FORMATS[clip["format"]]


def _plane_mask(clip, planes, default_all=True):
    n = len(clip["planes"])
    if planes is None:
        return [True] * n if default_all else [i == 0 for i in range(n)]
    return [i in planes for i in range(n)]


def boxblur(clip, planes=None, hradius=1, hpasses=1, vradius=1, vpasses=1):
    mask = _plane_mask(clip, planes)
    out = [oracle.boxblur_plane(p, hradius, hpasses, vradius, vpasses) if m else p.copy()
           for p, m in zip(clip["planes"], mask)]
    return {"format": clip["format"], "planes": out}


def _listify(v):
    if v is None:
        return []
    return list(v) if isinstance(v, (list, tuple)) else [v]


def bilateral(clip, ref=None, sigmaS=None, sigmaR=None, planes=None, algorithm=None, PBFICnum=None):
    fam, st, bits, ssw, ssh = _fmt(clip)
    rc, prm = oracle.bilateral_derive(fam == "YUV", st == "f", bits, ssw, ssh, len(clip["planes"]),
                                      _listify(sigmaS), _listify(sigmaR), planes, _listify(algorithm), _listify(PBFICnum))
    assert rc == 0, f"oracle bilateral_derive rc={rc}"
    out = []
    for i, p in enumerate(clip["planes"]):
        if not prm.process[i]:
            out.append(p.copy())
            continue
        r = None if ref is None else ref["planes"][i]
        if prm.algorithm[i] == 1:
            out.append(oracle.bilateral_pbfic_plane(p, prm.sigmaS[i], prm.sigmaR[i], prm.pbfic_num[i], prm.hist_len, r))
            continue
        out.append(oracle.bilateral_plane(p, prm.sigmaS[i], prm.sigmaR[i], prm.radius[i], prm.step[i], prm.hist_len, r))
    return {"format": clip["format"], "planes": out}


def limiter(clip, min=None, max=None, tv_range=False, mask=False, planes=None):
    """src/vapoursynth/limiter.zig:100-233 (valid arguments only) + the range tables of src/filters/limiter.zig:66-91."""
    fam, st, bits, ssw, ssh = _fmt(clip)
    pm = _plane_mask(clip, planes)
    n = len(clip["planes"])
    if min is not None:
        assert max is not None and len(min) == len(max) == n
        if st == "i":
            lo, hi = [float(np.trunc(v)) for v in min], [float(np.trunc(v)) for v in max]
        else:
            lo, hi = [float(np.float32(v)) for v in min], [float(np.float32(v)) for v in max]
    else:
        yuv = fam == "YUV" and not mask
        if st == "f":
            lo = [0.0] + [-0.5 if yuv else 0.0] * 2
            hi = [1.0] + [0.5 if yuv else 1.0] * 2
        elif tv_range:
            lo = [float(16 << (bits - 8))] * 3
            hi = [float(235 << (bits - 8))] + [float((240 if yuv else 235) << (bits - 8))] * 2
        else:
            lo, hi = [0.0] * 3, [float((1 << bits) - 1)] * 3
    out = [oracle.limiter_plane(p, lo[i], hi[i]) if m else p.copy() for i, (p, m) in enumerate(zip(clip["planes"], pm))]
    return {"format": clip["format"], "planes": out}


def scale_value(value, fam, st, bits, color_range=None):
    """hz.scaleValue(value, node, {depth_in=8, Integer, chroma=false}) (src/helper.zig:312-338), f32 arithmetic.
    color_range: 0 full / 1 limited as in frame 0's _ColorRange, None = prop absent (RGB -> full, else limited; helper.zig:259-276)."""
    f32 = np.float32
    v = f32(value)
    if bits == 8:
        return float(v)
    limited = (fam != "RGB") if color_range is None else (color_range == 1)
    in_peak, in_low = (f32(235), f32(16)) if limited else (f32(255), f32(0))
    if st == "f":
        out_peak, out_low = f32(1), f32(0)
    elif limited:
        out_peak, out_low = f32(235 << (bits - 8)), f32(16 << (bits - 8))
    else:
        out_peak, out_low = f32((1 << bits) - 1), f32(0)
    v = f32(v * f32(f32(out_peak - out_low) / f32(in_peak - in_low)))
    if st == "i":
        r = f32(np.floor(np.abs(v) + f32(0.5)) * np.sign(v))  # @round: half away from zero
        v = f32(max(min(r, f32((1 << bits) - 1)), f32(0)))
    return float(v)


def _array3(v, default):
    """hz.getArray (src/helper.zig:340-404): up to 3 values, missing entries repeat the previous one."""
    v = _listify(v)
    out = []
    for i in range(3):
        out.append(v[i] if i < len(v) else (default if i == 0 else out[i - 1]))
    return out


def limitfilter(flt, src, ref=None, dark_thr=None, bright_thr=None, elast=None, planes=None, color_range=None):
    """src/vapoursynth/limit_filter.zig:93-124 (valid arguments only) + src/filters/limit_filter.zig:3-34."""
    fam, st, bits, ssw, ssh = _fmt(flt)
    pm = _plane_mask(flt, planes)
    dk = [scale_value(v, fam, st, bits, color_range) for v in _array3(dark_thr, 1.0)]
    br = [scale_value(v, fam, st, bits, color_range) for v in _array3(bright_thr, 1.0)]
    el = [float(np.float32(v)) for v in _array3(elast, 2.0)]
    out = []
    for i, (p, m) in enumerate(zip(flt["planes"], pm)):
        if not m:
            out.append(p.copy())
            continue
        out.append(oracle.limitfilter_plane(p, src["planes"][i], None if ref is None else ref["planes"][i], dk[i], br[i], el[i]))
    return {"format": flt["format"], "planes": out}


def adaptive_binarize(clip, clip2, c=3):
    """src/vapoursynth/adaptive_binarize.zig:28-70,97-99: every plane, c clamped to [-256, 256]."""
    c = int(np.clip(c, -256, 256))
    return {"format": clip["format"], "planes": [oracle.adaptive_binarize_plane(a, b, c) for a, b in zip(clip["planes"], clip2["planes"])]}


def _props(values_per_plane, keys, prop):
    """Append semantics of the reference: scalar for one processed plane, list for several."""
    res = {}
    for k in keys:
        vals = [v[k] for v in values_per_plane if k in v]
        if vals:
            res[prop + k] = vals[0] if len(vals) == 1 else vals
    return res


def planeminmax(clip, minthr=0.0, maxthr=0.0, clipb=None, planes=None, prop="psm"):
    fam, st, bits, ssw, ssh = _fmt(clip)
    mask = _plane_mask(clip, planes, default_all=False)
    vals = []
    for i, (p, m) in enumerate(zip(clip["planes"], mask)):
        if m:
            vals.append(oracle.planeminmax_plane(p, bits, minthr, maxthr, None if clipb is None else clipb["planes"][i]))
    return _props(vals, ("Min", "Max", "Diff"), prop)


def planeaverage(clip, exclude, clipb=None, planes=None, prop="psm"):
    fam, st, bits, ssw, ssh = _fmt(clip)
    mask = _plane_mask(clip, planes, default_all=False)
    vals = []
    for i, (p, m) in enumerate(zip(clip["planes"], mask)):
        if m:
            vals.append(oracle.planeaverage_plane(p, bits, exclude, None if clipb is None else clipb["planes"][i]))
    return _props(vals, ("Avg", "Diff"), prop)


def golden_stats(clip):
    """tests/golden.py:106-121 of the reference: per-plane {avg,min,max} via std.PlaneStats."""
    fam, st, bits, ssw, ssh = _fmt(clip)
    return {f"p{i}": oracle.plane_stats(p, bits) for i, p in enumerate(clip["planes"])}


def parse_case_id(key: str):
    """Inverse of the reference's Case.id (tests/golden.py:40-56): 'FMT|geometry|k=v,...[|variant]'."""
    parts = key.split("|")
    fmt, geometry, argstr = parts[0], parts[1], parts[2]
    variant = parts[3] if len(parts) > 3 else ""
    args = {}
    if argstr != "default":
        # split on commas that are not inside [...]
        items, depth, cur = [], 0, ""
        for ch in argstr:
            if ch == "[":
                depth += 1
            elif ch == "]":
                depth -= 1
            if ch == "," and depth == 0:
                items.append(cur)
                cur = ""
            else:
                cur += ch
        items.append(cur)
        for it in items:
            k, v = it.split("=", 1)
            if v.startswith("["):
                args[k] = [_num(x) for x in v[1:-1].split(",") if x]
            else:
                args[k] = _num(v)
    return fmt, geometry, args, variant


def _num(s):
    try:
        return int(s)
    except ValueError:
        try:
            return float(s)
        except ValueError:
            return s
