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
