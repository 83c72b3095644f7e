"""LimitFilter and AdaptiveBinarize on the GPU against the CPU oracle: bit-exact on every sample type,
and against the reference's recorded goldens (tests/golden/limitfilter.json) / known-answer tests."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, from_frame, noise_clip, to_node
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "limitfilter.json").read_text())


def node(clip, color_range=None):
    props = None if color_range is None else {"_ColorRange": color_range}
    return vz.core.clip_from_frames(clip["format"], [clip["planes"]], props)


# --------------------------------------------------------------------------- LimitFilter
@pytest.mark.parametrize("key", sorted(k for k in GOLD if k.split("|")[0] in fx.FORMATS))
def test_limitfilter_golden_cases(key):
    """flt = BoxBlur(2,2), ref = BoxBlur(4,4) - both computed by the product too, like the reference's own chains."""
    fmt, geo, args, variant = oa.parse_case_id(key)
    src = fx.make_clip(fmt, geo)
    s = node(src, 0)   # the goldens were recorded with the thresholds scaled on the full-range rule (see test_oracle_goldens.py)
    flt = s.vszip.BoxBlur(hradius=2, vradius=2)
    ref = s.vszip.BoxBlur(hradius=4, vradius=4) if variant == "ref" else None
    got = from_frame(fmt, flt.vszip.LimitFilter(s, ref, **args).get_frame(0))
    o_flt = oa.boxblur(src, hradius=2, vradius=2)
    o_ref = oa.boxblur(src, hradius=4, vradius=4) if variant == "ref" else None
    assert_same_planes(got["planes"], oa.limitfilter(o_flt, src, o_ref, color_range=0, **args)["planes"], key)
    stats = oa.golden_stats(got)
    for p, e in GOLD[key].items():
        assert stats[p]["min"] == e["min"] and stats[p]["max"] == e["max"]
        assert stats[p]["avg"] == pytest.approx(e["avg"], rel=1e-6)  # the reference suite's own tolerance


def _near(clip, seed, spread):
    """A clip that differs from `clip` by small amounts, so that all three branches of the soft limit are taken."""
    fam, st, bits, ssw, ssh = fx.FORMATS[clip["format"]]
    rng = np.random.default_rng(seed)
    out = []
    for p in clip["planes"]:
        if st == "i":
            d = rng.integers(-spread, spread + 1, size=p.shape)
            out.append(np.clip(p.astype(np.int64) + d, 0, (1 << bits) - 1).astype(p.dtype))
        else:
            d = (rng.random(p.shape, dtype=np.float32) - np.float32(0.5)) * np.float32(2 * spread / 255.0)
            out.append((p.astype(np.float32) + d).astype(p.dtype))
    return {"format": clip["format"], "planes": out}


@pytest.mark.parametrize("fmt", ["GRAY8", "GRAY10", "GRAY16", "GRAYH", "GRAYS", "YUV420P8", "YUV420P16", "YUV444PS", "RGB24", "RGBS"])
def test_limitfilter_noise(fmt):
    base = "GRAY16" if fmt == "GRAY10" else fmt
    src = noise_clip(base, 517, 243, seed=71)   # odd width: vector body + scalar row tails
    if fmt == "GRAY10":
        src = {"format": "GRAY10", "planes": [src["planes"][0] >> 6]}
    fam, st, bits, ssw, ssh = fx.FORMATS[fmt]
    unit = 1 if bits == 8 or st == "f" else (1 << (bits - 8))
    flt, ref = _near(src, 72, 24 * unit if st == "i" else 24), _near(src, 73, 6 * unit if st == "i" else 6)
    n = len(src["planes"])
    cases = [dict(), dict(dark_thr=4, bright_thr=4, elast=2), dict(dark_thr=16, bright_thr=2, elast=4), dict(dark_thr=8, bright_thr=16, elast=1.5),
             dict(dark_thr=0, bright_thr=0, elast=0), dict(dark_thr=3, bright_thr=3, elast=1), dict(dark_thr=255, bright_thr=255, elast=65535)]
    if n == 3:
        cases += [dict(dark_thr=[16, 4], bright_thr=[16, 4, 9], elast=[4, 2]), dict(dark_thr=8, planes=[1]), dict(dark_thr=8, planes=[0, 2])]
    for cr in (None, 0, 1):
        for args in cases:
            for r, o_r in ((None, None), (node(ref, cr), ref)):
                got = from_frame(fmt, node(flt, cr).vszip.LimitFilter(node(src, cr), r, **args).get_frame(0))
                assert_same_planes(got["planes"], oa.limitfilter(flt, src, o_r, color_range=cr, **args)["planes"], f"{fmt} cr={cr} {args} ref={r is not None}")


def test_limitfilter_device_batch():
    fmt, w, h, n = "YUV420P16", 640, 360, 3
    a, b, c, d = (vz.DeviceClip(fmt, w, h, n) for _ in range(4))
    a.fill_noise(seed=5)
    vz.BoxBlurFilter(a.info(), hradius=2, vradius=2).run_device(a, b)
    vz.BoxBlurFilter(a.info(), hradius=4, vradius=4).run_device(a, c)
    for use_ref in (False, True):
        f = vz.LimitFilterFilter(a.info(), a.info(), a.info() if use_ref else None, dark_thr=8, bright_thr=4, elast=3, color_range=1)
        f.run_device(b, a, d, ref=c if use_ref else None)
        for i in range(n):
            src, flt, ref = ({"format": fmt, "planes": x.download(i)} for x in (a, b, c))
            want = oa.limitfilter(flt, src, ref if use_ref else None, dark_thr=8, bright_thr=4, elast=3, color_range=1)
            assert_same_planes(d.download(i), want["planes"], f"frame {i} ref={use_ref}")


# --------------------------------------------------------------------------- AdaptiveBinarize
@pytest.mark.parametrize("c", [0, 3, 10])
def test_adaptive_binarize_threshold_rule(c):
    """tests/test_adaptive_binarize.py:84-96 of the reference."""
    ramp = np.tile(np.arange(256, dtype=np.uint8), (2, 1))
    src = vz.core.clip_from_frames("GRAY8", [[ramp]])
    out = src.vszip.AdaptiveBinarize(vz.core.BlankClip("GRAY8", 256, 2, color=128), c=c).get_frame(0)
    assert out.planes[0][0].tolist() == [255 if x <= 128 - c else 0 for x in range(256)]
    assert out.props["_ColorRange"] == 0


@pytest.mark.parametrize("fmt", ["GRAY8", "YUV420P8", "RGB24"])
def test_adaptive_binarize_noise(fmt):
    a = noise_clip(fmt, 517, 243, seed=81)
    b = _near(a, 82, 20)
    for c in (None, -300, -256, -255, -254, -5, -1, 0, 1, 3, 12, 254, 255, 256, 300):
        args = {} if c is None else dict(c=c)
        got = from_frame(fmt, to_node(a).vszip.AdaptiveBinarize(to_node(b), **args).get_frame(0))
        assert_same_planes(got["planes"], oa.adaptive_binarize(a, b, **args)["planes"], f"{fmt} c={c}")
    # extreme differences: 0 vs 255 on both sides
    lo = {"format": fmt, "planes": [np.zeros_like(p) for p in a["planes"]]}
    hi = {"format": fmt, "planes": [np.full_like(p, 255) for p in a["planes"]]}
    for c in (-256, -255, 0, 255, 256):
        for x, y in ((lo, hi), (hi, lo)):
            got = from_frame(fmt, to_node(x).vszip.AdaptiveBinarize(to_node(y), c=c).get_frame(0))
            assert_same_planes(got["planes"], oa.adaptive_binarize(x, y, c=c)["planes"], f"{fmt} extremes c={c}")


def test_adaptive_binarize_with_vszip_blur_and_device_batch():
    """The reference's usage (clip2 = a blurred copy), with vszip's own BoxBlur; then the batched device entry point."""
    src = fx.make_clip("GRAY8", "full")
    s = to_node(src)
    got = from_frame("GRAY8", s.vszip.AdaptiveBinarize(s.vszip.BoxBlur(hradius=5, vradius=5)).get_frame(0))
    want = oa.adaptive_binarize(src, oa.boxblur(src, hradius=5, vradius=5))
    assert_same_planes(got["planes"], want["planes"], "GRAY8 fixture")
    assert set(np.unique(got["planes"][0]).tolist()) <= {0, 255}
    fmt, w, h, n = "YUV420P8", 640, 360, 3
    a, b, d = (vz.DeviceClip(fmt, w, h, n) for _ in range(3))
    a.fill_noise(seed=15)
    vz.BoxBlurFilter(a.info(), hradius=5, vradius=5).run_device(a, b)
    vz.AdaptiveBinarizeFilter(a.info(), a.info(), c=6).run_device(a, b, d)
    for i in range(n):
        x, y = ({"format": fmt, "planes": c.download(i)} for c in (a, b))
        assert_same_planes(d.download(i), oa.adaptive_binarize(x, y, c=6)["planes"], f"frame {i}")


# --------------------------------------------------------------------------- BASELINE-size frames, size-independent properties
def test_full_size_properties_1080p():
    """1920x1080 YUV420P16 / YUV420P8 device batches (too large for the scalar oracle to be the only check):
    LimitFilter(flt, flt) == flt;  thr = 0, elast = 0 returns src wherever flt != ref;  an infinite elast with thr = 255 on 8-bit-scale
    noise keeps flt wherever |flt - src| <= thr;  AdaptiveBinarize(a, a, c=0) is all 255, c=1 all 0;  one frame against the oracle."""
    fmt, w, h, n = "YUV420P16", 1920, 1080, 4
    a, b, d = (vz.DeviceClip(fmt, w, h, n) for _ in range(3))
    a.fill_noise(seed=31)
    b.fill_noise(seed=32)
    vi = a.info()
    vz.LimitFilterFilter(vi, vi, None, dark_thr=7, bright_thr=3, elast=2.5).run_device(a, a, d)
    for i in (0, n - 1):
        assert_same_planes(d.download(i), a.download(i), "LimitFilter(flt, flt) == flt")
    vz.LimitFilterFilter(vi, vi, None, dark_thr=0, bright_thr=0, elast=0).run_device(a, b, d)
    for i in (0, n - 1):
        assert_same_planes(d.download(i), b.download(i), "thr = 0: src everywhere (flt == src returns flt == src)")
    f = vz.LimitFilterFilter(vi, vi, None, dark_thr=16, bright_thr=4, elast=3, color_range=0)
    f.run_device(a, b, d)
    flt, src = ({"format": fmt, "planes": c.download(1)} for c in (a, b))
    assert_same_planes(d.download(1), oa.limitfilter(flt, src, None, dark_thr=16, bright_thr=4, elast=3, color_range=0)["planes"], "1080p frame vs oracle")
    fmt8 = "YUV420P8"
    a8, d8 = vz.DeviceClip(fmt8, w, h, n), vz.DeviceClip(fmt8, w, h, n)
    a8.fill_noise(seed=33)
    vz.AdaptiveBinarizeFilter(a8.info(), a8.info(), c=0).run_device(a8, a8, d8)
    assert all(int(p.min()) == 255 for p in d8.download(n - 1))
    vz.AdaptiveBinarizeFilter(a8.info(), a8.info(), c=1).run_device(a8, a8, d8)
    assert all(int(p.max()) == 0 for p in d8.download(0))
