"""BASELINE config 5: BoxBlur -> Bilateral -> PlaneMinMax chained on device-resident YUV444PS frames
(no PCIe round trips between the filters), against the same chain run through the CPU oracle."""
import numpy as np
import pytest

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, noise_clip, to_node

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize(("fmt", "w", "h"), [("YUV444PS", 480, 270), ("YUV420P16", 480, 270)])
def test_chain_device_resident(fmt, w, h):
    n = 3
    a, b, c = (vz.DeviceClip(fmt, w, h, n) for _ in range(3))
    a.fill_noise(seed=77, first_frame_no=0, frame_no_stride=2)    # the frames a rank-0-of-2 process would own
    blur = vz.BoxBlurFilter(a.info(), hradius=13, hpasses=1, vradius=13, vpasses=1)
    bil = vz.BilateralFilter(a.info(), sigmaS=2, sigmaR=2)
    mm = vz.PlaneMinMaxFilter(a.info(), minthr=0.1, maxthr=0.1, planes=[0])
    blur.run_device(a, b)
    bil.run_device(b, c)
    props = mm.run_device(c)
    is_float = vz.FORMATS[fmt].sample_type == vz.FLOAT
    for i in range(n):
        src = {"format": fmt, "planes": a.download(i)}
        want_blur = oa.boxblur(src, hradius=13, vradius=13)
        assert_same_planes(b.download(i), want_blur["planes"], f"frame {i} BoxBlur")          # bit-exact, also for f32
        want_bil = oa.bilateral(want_blur, sigmaS=2, sigmaR=2)
        got_bil = c.download(i)
        for g, wv in zip(got_bil, want_bil["planes"]):
            if is_float:
                assert np.all(np.abs(g.astype(np.float64) - wv) <= 1e-5 * np.abs(wv) + 1e-6)
            else:
                assert np.abs(g.astype(np.int64) - wv.astype(np.int64)).max() <= 1
        # the reduction is exact on whatever frame it is given: check it on the GPU's own bilateral output
        want_mm = oa.planeminmax({"format": fmt, "planes": got_bil}, minthr=0.1, maxthr=0.1, planes=[0])
        assert props[i] == want_mm, (props[i], want_mm)


# --------------------------------------------------------------------------- fused get_frame chains (vszip_chain_*)
def _chain_case(fmt, w, h, build):
    clip = noise_clip(fmt, w, h, seed=71)
    node = build(to_node(clip))
    vz.core.fuse_chains = True
    try:
        fused = node.get_frame(0)
        assert getattr(node, "_chain", None) is not None, "the chain was not fused"
    finally:
        vz.core.fuse_chains = False
    try:
        plain = build(to_node(clip)).get_frame(0)   # one upload/download per filter
    finally:
        vz.core.fuse_chains = True
    assert_same_planes(fused.planes, plain.planes, f"fused chain {fmt}")
    assert fused.props == plain.props, (fused.props, plain.props)
    return fused


def test_fused_chain_config5_shape():
    """BASELINE config 5's chain through the frame API: BoxBlur -> Bilateral -> PlaneMinMax, one PCIe round trip."""
    f = _chain_case("YUV444PS", 320, 180, lambda c: c.vszip.BoxBlur(hradius=13, vradius=13).vszip.Bilateral(sigmaS=2, sigmaR=2)
                    .vszip.PlaneMinMax(minthr=0.1, maxthr=0.1, planes=[0]))
    assert "psmMin" in f.props and "psmMax" in f.props


def test_fused_chain_partial_planes_and_stats_in_the_middle():
    def build(c):
        c = c.vszip.BoxBlur(planes=[0], hradius=3, hpasses=2, vradius=0, vpasses=0)
        c = c.vszip.PlaneAverage(exclude=[0, 7], planes=[0, 1, 2], prop="a")
        c = c.vszip.Bilateral(sigmaS=1.5, sigmaR=0.02, planes=[1, 2])
        c = c.vszip.PlaneMinMax(minthr=0.05, maxthr=0.2, planes=[0, 2], prop="b")
        return c.vszip.BoxBlur(planes=[2], hradius=0, hpasses=0, vradius=5, vpasses=3)
    f = _chain_case("YUV420P16", 322, 182, build)
    assert isinstance(f.props["aAvg"], list) and len(f.props["bMin"]) == 2


def test_fused_chain_stats_only_and_pbfic():
    _chain_case("GRAY16", 200, 120, lambda c: c.vszip.PlaneMinMax(minthr=0.1).vszip.PlaneAverage(exclude=[5]))
    _chain_case("GRAY16", 200, 120, lambda c: c.vszip.Bilateral(sigmaS=3, sigmaR=0.1, algorithm=1).vszip.BoxBlur(hradius=2, vradius=2))


def test_chain_rejects_second_clips():
    a = to_node(noise_clip("GRAY8", 64, 48, seed=1))
    b = to_node(noise_clip("GRAY8", 64, 48, seed=2))
    n = a.vszip.BoxBlur().vszip.PlaneMinMax(clipb=b)
    out = n.get_frame(0)                      # evaluated unfused: the second clip keeps it out of a chain
    assert getattr(n, "_chain", None) is None and "psmDiff" in out.props


# --------------------------------------------------------------------------- diamond elements: the second clip is the chain's source
def test_fused_chain_limitfilter_reads_the_chain_source():
    """`flt = src.vszip.BoxBlur(); flt.vszip.LimitFilter(src)` (the reference's usage, tests/test_int_parity.py:158-167) as one chain:
    one upload, one download; identical to the two separate calls and to the oracle."""
    def build(c):
        return c.vszip.BoxBlur(hradius=2, vradius=2).vszip.LimitFilter(c, dark_thr=[16, 4], bright_thr=[8, 4], elast=[3, 2], planes=[0, 2]) \
                .vszip.PlaneMinMax(minthr=0.05, maxthr=0.05)
    f = _chain_case("YUV420P16", 322, 182, build)
    clip = noise_clip("YUV420P16", 322, 182, seed=71)
    want = oa.limitfilter(oa.boxblur(clip, hradius=2, vradius=2), clip, None, dark_thr=[16, 4], bright_thr=[8, 4], elast=[3, 2], planes=[0, 2])
    assert_same_planes(f.planes, want["planes"], "BoxBlur -> LimitFilter(src) chain vs oracle")
    # three pixel filters: the source must survive the ping-pong of the intermediates
    _chain_case("GRAYS", 200, 120, lambda c: c.vszip.BoxBlur(hradius=1, vradius=1).vszip.BoxBlur(hradius=3, vradius=0, vpasses=0)
                .vszip.Limiter(min=[0.1], max=[0.9]).vszip.LimitFilter(c, dark_thr=8, bright_thr=8, elast=1.5))


def test_fused_chain_adaptive_binarize():
    """`src.vszip.AdaptiveBinarize(src.vszip.BoxBlur(5, 5))` (tests/test_adaptive_binarize.py:59-63 with vszip's own blur)."""
    for fmt in ("GRAY8", "YUV420P8"):
        f = _chain_case(fmt, 322, 182, lambda c: c.vszip.AdaptiveBinarize(c.vszip.BoxBlur(hradius=5, vradius=5), c=2))
        clip = noise_clip(fmt, 322, 182, seed=71)
        assert_same_planes(f.planes, oa.adaptive_binarize(clip, oa.boxblur(clip, hradius=5, vradius=5), c=2)["planes"], fmt)
        assert f.props["_ColorRange"] == 0


def test_diamond_needs_the_chain_source():
    """A LimitFilter whose src is NOT the chain's source, or that has a ref clip, is evaluated unfused (and still correct)."""
    a = noise_clip("GRAY16", 200, 120, seed=1)
    b = noise_clip("GRAY16", 200, 120, seed=2)
    na, nb = to_node(a), to_node(b)
    flt = na.vszip.BoxBlur(hradius=2, vradius=2)
    n1 = flt.vszip.LimitFilter(nb, dark_thr=8)
    out1 = n1.get_frame(0)
    assert getattr(n1, "_chain", None) is None
    assert_same_planes(out1.planes, oa.limitfilter(oa.boxblur(a, hradius=2, vradius=2), b, None, dark_thr=8)["planes"], "foreign src")
    n2 = flt.vszip.LimitFilter(na, nb, dark_thr=8)
    out2 = n2.get_frame(0)
    assert getattr(n2, "_chain", None) is None
    assert_same_planes(out2.planes, oa.limitfilter(oa.boxblur(a, hradius=2, vradius=2), a, b, dark_thr=8)["planes"], "ref clip")


def test_fused_chain_props_follow_the_first_clip():
    """AdaptiveBinarize allocates dst from `clip`, so stats props computed on the clip2 branch must not leak into the result,
    fused or not (src/vapoursynth/adaptive_binarize.zig:28-60)."""
    f = _chain_case("GRAY8", 200, 120, lambda c: c.vszip.AdaptiveBinarize(c.vszip.BoxBlur(hradius=3, vradius=3).vszip.PlaneMinMax(minthr=0.1), c=1))
    assert "psmMin" not in f.props and f.props["_ColorRange"] == 0
