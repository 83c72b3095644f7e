"""Pins the CPU oracle to the reference's own golden vectors.

tests/golden/*.json are verbatim copies of /root/reference/tests/goldens/*.json (values recorded
from the real Zig plugin under VapourSynth).  Each case rebuilds the reference's fixture clip
(oracle/fixtures.py), runs the oracle and compares with the recorded snapshot - tighter than the
reference's own rel=1e-6 (tests/golden.py:187-191): averages to 1e-12 relative, min/max exact.

Keys that need VapourSynth's std.BoxBlur (a different plugin) for their second clip are not
reproducible here and are listed as skipped.
"""
import json
from pathlib import Path

import pytest

import oracle_api as oa
from oracle import fixtures as fx

GOLD = Path(__file__).resolve().parent / "golden"


def _load(name):
    return json.loads((GOLD / f"{name}.json").read_text())


def _close(a, e, rel):
    return abs(a - e) <= rel * max(abs(e), 1e-300) or abs(a - e) <= 1e-15


def _assert_stats(actual, expected, rel=1e-12):
    assert set(actual) == set(expected)
    for p in expected:
        assert actual[p]["min"] == expected[p]["min"], (p, actual[p], expected[p])
        assert actual[p]["max"] == expected[p]["max"], (p, actual[p], expected[p])
        assert _close(actual[p]["avg"], expected[p]["avg"], rel), (p, actual[p], expected[p])


def _assert_props(actual, expected, rel=1e-12):
    assert set(actual) == set(expected), (actual, expected)
    for k, e in expected.items():
        a = actual[k]
        if isinstance(e, list):
            assert isinstance(a, list) and len(a) == len(e), (k, a, e)
            for x, y in zip(a, e):
                assert _close(x, y, rel), (k, a, e)
        else:
            assert _close(a, e, rel), (k, a, e)


# --------------------------------------------------------------------------- BoxBlur
@pytest.mark.parametrize("key", sorted(_load("boxblur")))
def test_boxblur_golden(key):
    fmt, geo, args, variant = oa.parse_case_id(key)
    out = oa.boxblur(fx.make_clip(fmt, geo), **args)
    _assert_stats(oa.golden_stats(out), _load("boxblur")[key])


# --------------------------------------------------------------------------- Bilateral
@pytest.mark.parametrize("key", sorted(_load("bilateral")))
def test_bilateral_golden(key):
    fmt, geo, args, variant = oa.parse_case_id(key)
    if variant == "ref":
        pytest.skip("joint clip is built with VapourSynth's std.BoxBlur (not part of vszip)")
    out = oa.bilateral(fx.make_clip(fmt, geo), **args)
    _assert_stats(oa.golden_stats(out), _load("bilateral")[key])


# --------------------------------------------------------------------------- Limiter
@pytest.mark.parametrize("key", sorted(_load("limiter")))
def test_limiter_golden(key):
    fmt, geo, args, variant = oa.parse_case_id(key)
    if fmt not in fx.FORMATS:
        pytest.skip(f"the fixture generator does not restate zimg's conversion to {fmt}")
    out = oa.limiter(fx.make_clip(fmt, geo), **args)
    # the restated zimg chain reproduces subsampled f32 chroma to ~1e-7 per sample only (plane averages to ~2e-11):
    # YUV420PS keys use 1e-9 here, still far inside the reference suite's own 1e-6
    _assert_stats(oa.golden_stats(out), _load("limiter")[key], rel=1e-9 if fmt == "YUV420PS" else 1e-12)


# --------------------------------------------------------------------------- LimitFilter
# The reference tree holds tests/goldens/limitfilter.json but no longer the test that recorded it; the construction is the one its
# parity tests still use (tests/test_int_parity.py:158-167, tests/test_f16_parity.py:211-244): flt = src.vszip.BoxBlur(2, 2),
# ref = src.vszip.BoxBlur(4, 4) for the "ref" variant.  Every reproducible key matches exactly with the thresholds scaled on the
# FULL-range rule (value * peak / 255), i.e. hz.getColorRange returned FULL for the fixture clips when the goldens were recorded.
@pytest.mark.parametrize("key", sorted(_load("limitfilter")))
def test_limitfilter_golden(key):
    fmt, geo, args, variant = oa.parse_case_id(key)
    if fmt not in fx.FORMATS:
        pytest.skip(f"the fixture generator does not restate zimg's conversion to {fmt}")
    src = fx.make_clip(fmt, geo)
    flt = oa.boxblur(src, hradius=2, vradius=2)
    ref = oa.boxblur(src, hradius=4, vradius=4) if variant == "ref" else None
    out = oa.limitfilter(flt, src, ref, color_range=0, **args)
    _assert_stats(oa.golden_stats(out), _load("limitfilter")[key])


# --------------------------------------------------------------------------- AdaptiveBinarize
@pytest.mark.parametrize("key", sorted(_load("adaptive_binarize")))
def test_adaptive_binarize_golden(key):
    pytest.skip("clip2 is built with VapourSynth's std.BoxBlur (not part of vszip); pinned by the known-answer tests below")


@pytest.mark.parametrize("c", [0, 3, 10])
def test_adaptive_binarize_threshold_rule(c):
    """tests/test_adaptive_binarize.py:84-96 of the reference: a 0..255 ramp against a constant 128."""
    import numpy as np
    ramp = np.tile(np.arange(256, dtype=np.uint8), (2, 1))
    out = oa.adaptive_binarize({"format": "GRAY8", "planes": [ramp]}, {"format": "GRAY8", "planes": [np.full_like(ramp, 128)]}, c=c)
    assert out["planes"][0][0].tolist() == [255 if x <= 128 - c else 0 for x in range(256)]


def test_adaptive_binarize_vszip_blur_mean():
    """tests/test_adaptive_binarize.py:99-103 (higher c is stricter), with vszip's own BoxBlur as the companion clip."""
    src = fx.make_clip("GRAY8", "full")
    blur = oa.boxblur(src, hradius=5, vradius=5)
    a3 = oa.adaptive_binarize(src, blur, c=3)["planes"][0].mean()
    a10 = oa.adaptive_binarize(src, blur, c=10)["planes"][0].mean()
    assert a10 < a3 and set(np_unique(oa.adaptive_binarize(src, blur)["planes"][0])) <= {0, 255}


def np_unique(a):
    import numpy as np
    return np.unique(a).tolist()


# --------------------------------------------------------------------------- PlaneMinMax
@pytest.mark.parametrize("key", sorted(_load("planeminmax")))
def test_planeminmax_golden(key):
    fmt, geo, args, variant = oa.parse_case_id(key)
    clip = fx.make_clip(fmt, geo)
    use_clipb = bool(args.pop("variant_clipb", 0)) or variant == "ref"
    if use_clipb:  # tests/test_planeminmax.py:71-72 of the reference: clipb = src.vszip.BoxBlur(1, 1)
        args["clipb"] = oa.boxblur(clip, hradius=1, vradius=1)
    prop = args.get("prop", "psm")
    got = oa.planeminmax(clip, **args)
    got = {k[len(prop):]: v for k, v in got.items()}
    _assert_props(got, _load("planeminmax")[key])


# --------------------------------------------------------------------------- PlaneAverage
@pytest.mark.parametrize("key", sorted(_load("planeaverage")))
def test_planeaverage_golden(key):
    fmt, geo, args, variant = oa.parse_case_id(key)
    if variant.startswith("ref"):
        pytest.skip("clipb is built with VapourSynth's std.BoxBlur (not part of vszip)")
    clip = fx.make_clip(fmt, geo)
    prop = args.get("prop", "psm")
    got = oa.planeaverage(clip, **args)
    got = {"avg": got[prop + "Avg"]}
    _assert_props(got, _load("planeaverage")[key])
