"""CPU-only: create-time argument handling of the four filters through the C ABI - defaults,
dispatch parameters and the reference's exact error substrings (tests/test_boxblur.py:131-163,
test_bilateral.py:102-164, test_planeminmax.py:163-224, test_planeaverage.py:150-218 of the
reference).  The Bilateral parameter derivation is checked against the oracle's restatement."""
import pytest

import oracle
import vapoursynth_zip_b200 as vz

core = vz.core


def raises(msg, fn):
    with pytest.raises(vz.Error, match=msg):
        fn()


# --------------------------------------------------------------------------- BoxBlur
@pytest.mark.parametrize(("args", "msg"), [
    (dict(hradius=0, vradius=0, hpasses=0, vpasses=0), "nothing to be performed"),
    (dict(hradius=5, vradius=5, hpasses=0, vpasses=0), "nothing to be performed"),
    (dict(planes=[3]), "plane index out of range"),
    (dict(planes=[-1]), "plane index out of range"),
    (dict(planes=[0, 0]), "plane specified twice"),
    (dict(hradius=16, vradius=1), "hradius too large; 2\\*hradius must be < the \\(smallest processed\\) plane width"),
    (dict(hradius=1, vradius=8), "vradius too large; 2\\*vradius must be < the \\(smallest processed\\) plane height"),
])
def test_boxblur_validation(args, msg):
    raises(msg, lambda: core.BlankClip("YUV420P8", 64, 32).vszip.BoxBlur(**args))


def test_boxblur_radius_check_only_on_processed_planes():
    # luma 64x32: hradius=20 is fine for plane 0, too large for the 32-wide chroma planes
    core.BlankClip("YUV420P8", 64, 32).vszip.BoxBlur(planes=[0], hradius=20, vradius=1)
    raises("hradius too large", lambda: core.BlankClip("YUV420P8", 64, 32).vszip.BoxBlur(hradius=20, vradius=1))


def test_boxblur_unsupported_format():
    raises("not supported Int format", lambda: core.BlankClip("GRAY32", 64, 64).vszip.BoxBlur(hradius=1, vradius=1))


def test_boxblur_default_planes_and_mask():
    node = core.BlankClip("YUV420P16", 64, 32).vszip.BoxBlur()
    assert node.filter.process == [True, True, True]
    assert core.BlankClip("YUV420P16", 64, 32).vszip.BoxBlur(planes=[1, 2]).filter.process == [False, True, True]
    assert core.BlankClip("GRAY8", 64, 32).vszip.BoxBlur().filter.process == [True, False, False]


# --------------------------------------------------------------------------- Bilateral
@pytest.mark.parametrize(("args", "msg"), [
    (dict(sigmaS=-1), 'Invalid "sigmaS" assigned'),
    (dict(PBFICnum=1), 'Invalid "PBFICnum" assigned'),
    (dict(PBFICnum=300), "PBFICnum value 300 is above maximum 256"),
    (dict(algorithm=3), "algorithm value 3 is above maximum 2"),
    (dict(sigmaR=-0.5), "sigmaR value -0.5 is below minimum 0"),
    (dict(sigmaR=[1, 2, 3, 4]), "sigmaR has too many elements \\(got 4, max 3\\)"),
    (dict(planes=[1]), "plane index out of range"),
])
def test_bilateral_validation(args, msg):
    raises(msg, lambda: core.BlankClip("GRAY16", 64, 64).vszip.Bilateral(**args))


@pytest.mark.parametrize(("w", "h"), [(20, 4), (5, 20), (4, 4), (3, 30)])
@pytest.mark.parametrize("fmt", ["GRAY8", "GRAY16", "GRAYS"])
def test_bilateral_small_frame_errors(fmt, w, h):
    raises("plane too small for the spatial radius", lambda: core.BlankClip(fmt, w, h).vszip.Bilateral())


def test_bilateral_small_chroma_errors():
    raises("plane too small for the spatial radius",
           lambda: core.BlankClip("YUV420P8", 64, 64).vszip.Bilateral(sigmaS=[2, 20], algorithm=2))


def test_bilateral_algorithm1_is_accepted():
    """PBFIC has a CUDA path; a plane on algorithm 1 is exempt from the radius size check (bilateral.zig:201-214)."""
    info = core.BlankClip("GRAY16", 8, 8).vszip.Bilateral(sigmaS=20, sigmaR=0.1, algorithm=1).filter.info()
    assert info.algorithm[0] == 1 and info.PBFICnum[0] == 4


def test_bilateral_ref_mismatch():
    a = core.BlankClip("YUV420P16", 64, 32, 3)
    raises("same width and height", lambda: a.vszip.Bilateral(ref=core.BlankClip("YUV420P16", 32, 32, 3)))
    raises("same bit depth", lambda: a.vszip.Bilateral(ref=core.BlankClip("YUV420P8", 64, 32, 3)))


CASES = [
    ("GRAY16", dict()), ("GRAY16", dict(sigmaS=2, sigmaR=2)), ("GRAY16", dict(sigmaS=0.8, sigmaR=0.02)),
    ("GRAY16", dict(sigmaS=5, sigmaR=0.02)), ("GRAY16", dict(sigmaS=5, sigmaR=2)), ("GRAY16", dict(sigmaS=3, sigmaR=0.02, algorithm=2)),
    ("YUV420P16", dict(sigmaS=2, sigmaR=2)), ("YUV420P16", dict(sigmaS=[3, 1.5], sigmaR=[0.02, 0.05])),
    ("YUV444P16", dict(sigmaS=2, sigmaR=2)), ("YUV420P8", dict(sigmaS=2, sigmaR=2, planes=[0])),
    ("RGBS", dict(sigmaS=2, sigmaR=2)), ("GRAY16", dict(sigmaS=0)), ("GRAY16", dict(sigmaS=10, sigmaR=0.02, algorithm=2)),
    ("GRAY8", dict(sigmaS=[1, 2, 4], sigmaR=[0.01, 0.1, 1.0], algorithm=2)),
]


@pytest.mark.parametrize(("fmt", "args"), CASES, ids=str)
def test_bilateral_derivation_matches_oracle(fmt, args):
    f = vz.FORMATS[fmt]
    info = core.BlankClip(fmt, 640, 320).vszip.Bilateral(**args).filter.info()
    as_list = lambda v: [] if v is None else (list(v) if isinstance(v, (list, tuple)) else [v])
    rc, want = oracle.bilateral_derive(f.color_family == vz.YUV, f.sample_type == vz.FLOAT, f.bits_per_sample, f.subsampling_w,
                                       f.subsampling_h, f.num_planes, as_list(args.get("sigmaS")), as_list(args.get("sigmaR")),
                                       args.get("planes"), as_list(args.get("algorithm")), as_list(args.get("PBFICnum")))
    assert rc == 0
    for i in range(f.num_planes):
        assert bool(info.process[i]) == bool(want.process[i])
        assert info.sigmaS[i] == want.sigmaS[i] and info.sigmaR[i] == want.sigmaR[i]
        if want.process[i]:
            assert (info.algorithm[i], info.PBFICnum[i], info.radius[i], info.samples[i], info.step[i]) == \
                   (want.algorithm[i], want.pbfic_num[i], want.radius[i], want.samples[i], want.step[i])


def test_config3_parameters():
    """luma r=3 step=2, chroma sigmaS=1 r=2 step=1, PBFICnum [4,5,5], algorithm 2 everywhere."""
    info = core.BlankClip("YUV420P16", 1920, 1080).vszip.Bilateral(sigmaS=2, sigmaR=2, planes=[0, 1, 2]).filter.info()
    assert list(info.sigmaS) == [2.0, 1.0, 1.0]
    assert (list(info.radius), list(info.step), list(info.algorithm), list(info.PBFICnum)) == ([3, 2, 2], [2, 1, 1], [2, 2, 2], [4, 5, 5])


# --------------------------------------------------------------------------- PlaneMinMax
@pytest.mark.parametrize(("args", "msg"), [
    (dict(minthr=1.5), "minthr should be a float between 0.0 and 1.0"),
    (dict(minthr=-0.1), "minthr should be a float between 0.0 and 1.0"),
    (dict(maxthr=2.0), "maxthr should be a float between 0.0 and 1.0"),
    (dict(maxthr=-0.5), "maxthr should be a float between 0.0 and 1.0"),
    (dict(planes=[3]), "plane index out of range"),
    (dict(planes=[-1]), "plane index out of range"),
    (dict(planes=[0, 0]), "plane specified twice"),
])
def test_planeminmax_validation(args, msg):
    raises(msg, lambda: core.BlankClip("YUV420P16", 64, 32).vszip.PlaneMinMax(**args))


def test_planeminmax_float_chroma_thr_error():
    raises("you can't use maxthr/minthr with float chroma",
           lambda: core.BlankClip("YUV420PS", 64, 32).vszip.PlaneMinMax(minthr=0.2, maxthr=0.3, planes=[0, 1, 2]))
    core.BlankClip("YUV420PS", 64, 32).vszip.PlaneMinMax(minthr=0.2, planes=[0])
    core.BlankClip("YUV420PS", 64, 32).vszip.PlaneMinMax(planes=[0, 1, 2])
    core.BlankClip("RGBS", 64, 32).vszip.PlaneMinMax(minthr=0.2, maxthr=0.3, planes=[0, 1, 2])


def test_planeminmax_int32_rejected():
    raises("not supported Int format", lambda: core.BlankClip("GRAY32", 64, 32).vszip.PlaneMinMax())


@pytest.mark.parametrize(("fmt", "dims", "msg"), [
    ("YUV420P16", (32, 32), "all input clips must have the same width and height"),
    ("RGB48", (64, 32), "all input clips must have the same color family"),
    ("YUV444P16", (64, 32), "all input clips must have the same subsampling"),
    ("YUV420P8", (64, 32), "all input clips must have the same bit depth"),
])
def test_clipb_mismatch_errors(fmt, dims, msg):
    a = core.BlankClip("YUV420P16", 64, 32, 3)
    b = core.BlankClip(fmt, dims[0], dims[1], 3)
    raises(msg, lambda: core.vszip.PlaneMinMax(clipa=a, clipb=b))
    raises(msg, lambda: core.vszip.PlaneAverage(clipa=a, exclude=[-1], clipb=b))


def test_clipb_shorter_error():
    a, b = core.BlankClip("GRAY8", 64, 32, 5), core.BlankClip("GRAY8", 64, 32, 3)
    raises("second clip has less frames than input clip", lambda: core.vszip.PlaneMinMax(clipa=a, clipb=b))
    raises("second clip has less frames than input clip", lambda: core.vszip.PlaneAverage(clipa=a, exclude=[-1], clipb=b))


def test_stats_default_plane_mask():
    assert core.BlankClip("YUV420P16", 64, 32).vszip.PlaneMinMax().filter.process == [True, False, False]
    assert core.BlankClip("YUV420P16", 64, 32).vszip.PlaneAverage(exclude=[-1], planes=[0, 2]).filter.process == [True, False, True]


# --------------------------------------------------------------------------- PlaneAverage
def test_planeaverage_exclude_required_and_int32():
    raises("exclude", lambda: core.BlankClip("GRAY16", 64, 32).vszip.PlaneAverage())
    raises("32-bit integer", lambda: core.BlankClip("GRAY32", 64, 32).vszip.PlaneAverage(exclude=[-1]))
    raises("plane index out of range", lambda: core.BlankClip("YUV420P16", 64, 32).vszip.PlaneAverage(exclude=[-1], planes=[3]))
    raises("plane specified twice", lambda: core.BlankClip("YUV420P16", 64, 32).vszip.PlaneAverage(exclude=[-1], planes=[0, 0]))


# --------------------------------------------------------------------------- Limiter (src/vapoursynth/limiter.zig:100-218)
@pytest.mark.parametrize(("args", "msg"), [
    (dict(min=[1, 2]), "min array must have the same number of elements as planes"),
    (dict(min=[-1, 0, 0], max=[1, 1, 1]), "min value must be greater than or equal to 0"),
    (dict(min=[70000, 0, 0], max=[1, 1, 1]), "min value must be less than or equal to peak value"),
    (dict(min=[0, 0, 0], max=[1, 1]), "max array must have the same number of elements as planes"),
    (dict(min=[0, 0, 0], max=[1, 70000, 1]), "max value must be less than or equal to peak value"),
    (dict(min=[0, 0, 0], max=[1, -3, 1]), "max value must be greater than or equal to 0"),
    (dict(min=[0, 0, 0]), "min array is set but max array is not"),
    (dict(max=[9, 9, 9]), "max array is set but min array is not"),
    (dict(min=[5, 0, 0], max=[4, 9, 9]), "min value must be less than or equal to max value"),
    (dict(planes=[3]), "plane index out of range"),
    (dict(planes=[1, 1]), "plane specified twice"),
])
def test_limiter_argument_errors(args, msg):
    raises("Limiter: " + msg, lambda: core.BlankClip("YUV444P16", 64, 32).vszip.Limiter(**args))


def test_limiter_format_errors_come_last():
    raises("Limiter: not supported Int format", lambda: core.BlankClip("GRAY11", 64, 32).vszip.Limiter())
    # a bad min array is reported before the format (limiter.zig:121 vs :220)
    raises("min array must have", lambda: core.BlankClip("GRAY11", 64, 32).vszip.Limiter(min=[1, 2]))


def test_limiter_accepts_32_bit_integer_clips():
    """BPSType.U32 (src/helper.zig:14-56, src/filters/limiter.zig:16,34,52): the Limiter is the one filter of the path that takes
    32-bit integer clips; its f32 peak is 2^32, so 4294967295 and even 4294967296 pass the peak check, 4294967808 does not."""
    c = core.BlankClip("GRAY32", 64, 32)
    for args in (dict(), dict(tv_range=True), dict(min=[7], max=[4294967295]), dict(min=[0], max=[4294967296])):
        assert c.vszip.Limiter(**args).format.bits_per_sample == 32
    raises("Limiter: max value must be less than or equal to peak value", lambda: c.vszip.Limiter(min=[0], max=[4294967808]))


# --------------------------------------------------------------------------- LimitFilter (src/vapoursynth/limit_filter.zig:93-124)
@pytest.mark.parametrize(("args", "msg"), [
    (dict(dark_thr=[1, 2, 3, 4]), "dark_thr has too many elements \\(got 4, max 3\\)"),
    (dict(dark_thr=-1), "dark_thr value -1 is below minimum 0"),
    (dict(bright_thr=300.5), "bright_thr value 300.5 is above maximum 255"),
    (dict(elast=70000), "elast value 70000 is above maximum 65535"),
    (dict(elast=[2, -0.25]), "elast value -0.25 is below minimum 0"),
    (dict(planes=[3]), "plane index out of range"),
    (dict(planes=[1, 1]), "plane specified twice"),
])
def test_limitfilter_argument_errors(args, msg):
    c = core.BlankClip("YUV444P16", 64, 32)
    raises("LimitFilter: " + msg, lambda: c.vszip.LimitFilter(c, **args))


@pytest.mark.parametrize(("other", "msg"), [
    (("GRAY16", 62, 32, 1), "same width and height"),
    (("YUV444P16", 64, 32, 1), "same color family"),
    (("GRAY8", 64, 32, 1), "same bit depth"),
    (("GRAY16", 64, 32, 2), "all input clips must have the same length"),
])
def test_limitfilter_clip_mismatch(other, msg):
    flt = core.BlankClip("GRAY16", 64, 32)
    fmt, w, h, n = other
    bad = core.BlankClip(fmt, w, h, n)
    raises("LimitFilter: .*" + msg, lambda: flt.vszip.LimitFilter(bad))
    raises("LimitFilter: .*" + msg, lambda: flt.vszip.LimitFilter(flt, bad))
    # DataType.select comes first (limit_filter.zig:101 vs :107)
    raises("LimitFilter: not supported Int format", lambda: core.BlankClip("GRAY32", 64, 32).vszip.LimitFilter(flt))


@pytest.mark.parametrize(("fmt", "cr", "want"), [
    ("GRAY8", None, (4.0, 8.0)),               # depth_in == depth_out: untouched
    ("GRAY16", 0, (1028.0, 2056.0)),           # full: 65535 / 255 = 257
    ("GRAY16", 1, (1024.0, 2048.0)),           # limited: (60160 - 4096) / 219 = 256
    ("GRAY16", None, (1024.0, 2048.0)),        # no _ColorRange: YUV/GRAY default to limited ...
    ("RGB48", None, (1028.0, 2056.0)),         # ... RGB to full (src/helper.zig:259-276)
    ("GRAY10", 0, (16.0, 32.0)),               # round(4 * 1023 / 255) = round(16.047)
    ("GRAYS", 0, (4 / 255, 8 / 255)),
    ("GRAYH", 1, (4 / 219, 8 / 219)),
])
def test_limitfilter_threshold_scaling(fmt, cr, want):
    import numpy as np
    if fmt not in ("GRAY8", "GRAY16", "GRAY10", "GRAYS", "GRAYH", "RGB48"):
        pytest.skip(fmt)
    try:
        c = core.BlankClip(fmt, 64, 32)
    except Exception:
        pytest.skip(f"{fmt} is not in the mirror's format table")
    f = vz.LimitFilterFilter(c._info(), c._info(), None, dark_thr=4, bright_thr=8, elast=[3, 1.5], color_range=cr)
    info = f.info()
    assert info["dark_thr"][0] == pytest.approx(want[0], rel=1e-6) and info["bright_thr"][2] == pytest.approx(want[1], rel=1e-6)
    assert info["elast"] == [3.0, 1.5, 1.5]
    # the same numbers as the oracle-side restatement of hz.scaleValue
    import oracle_api as oa
    from oracle.fixtures import FORMATS
    if fmt in FORMATS:
        fam, st, bits, _, _ = FORMATS[fmt]
        assert np.float32(info["dark_thr"][1]) == np.float32(oa.scale_value(4, fam, st, bits, cr))
        assert np.float32(info["bright_thr"][0]) == np.float32(oa.scale_value(8, fam, st, bits, cr))


# --------------------------------------------------------------------------- AdaptiveBinarize (adaptive_binarize.zig:79-116)
@pytest.mark.parametrize(("a", "b", "msg"), [
    (("GRAY16", 64, 32, 1), ("GRAY16", 64, 32, 1), "only 8 bit int format supported"),
    (("GRAYS", 64, 32, 1), ("GRAYS", 64, 32, 1), "only 8 bit int format supported"),
    (("GRAY8", 64, 32, 1), ("GRAY8", 62, 32, 1), "all input clips must have the same width and height"),
    (("GRAY8", 64, 32, 1), ("YUV444P8", 64, 32, 1), "all input clips must have the same color family"),
    (("YUV420P8", 64, 32, 1), ("YUV444P8", 64, 32, 1), "all input clips must have the same subsampling"),
    (("GRAY8", 64, 32, 1), ("GRAY16", 64, 32, 1), "all input clips must have the same bit depth"),
    (("GRAY8", 64, 32, 2), ("GRAY8", 64, 32, 1), "second clip has less frames than input clip"),
])
def test_adaptive_binarize_validation(a, b, msg):
    """tests/test_adaptive_binarize.py:112-134 of the reference."""
    ca, cb = core.BlankClip(*a), core.BlankClip(*b)
    raises("AdaptiveBinarize: " + msg, lambda: ca.vszip.AdaptiveBinarize(cb))
