"""fmParallel re-entrancy (src/vapoursynth/boxblur.zig:211, bilateral.zig:251, planeminmax.zig:170, planeaverage.zig:153 register
their filters as .Parallel): VapourSynth calls getFrame of ONE filter instance from all its worker threads at once, each for a
different frame.  Every filter handle is driven from 16 host threads here, more requests than the runtime has slots per GPU, and
every output frame / prop set is compared with the oracle (not with a serial GPU run, which could share a bug)."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, noise_clip

pytestmark = pytest.mark.gpu
THREADS, FRAMES, ROUNDS = 16, 24, 3


def _clips(fmt, w, h, seed0=900):
    return [noise_clip(fmt, w, h, seed=seed0 + i) for i in range(FRAMES)]


def _hammer(node, check):
    """every frame ROUNDS times, 16 requests in flight, in an order that makes neighbouring requests hit different frames"""
    order = [(i * 7) % FRAMES for i in range(FRAMES * ROUNDS)]
    with ThreadPoolExecutor(THREADS) as ex:
        results = list(ex.map(lambda n: (n, node.get_frame(n)), order))
    for n, fr in results:
        check(n, fr)


@pytest.mark.parametrize(("fmt", "args"), [
    ("YUV420P16", dict(hradius=13, hpasses=5, vradius=13, vpasses=5)),      # segment kernels (TMA-staged), config 2
    ("YUV420P16", dict(hradius=13, vradius=13)),                            # fused comptime kernel, config 1
    ("YUV420P8", dict(hradius=3, hpasses=2, vradius=5, vpasses=3)),         # streaming ring kernels
    ("GRAYS", dict(hradius=4, vradius=4)),                                  # comptime float
    ("GRAYH", dict(hradius=2, hpasses=3, vradius=0, vpasses=0)),
], ids=str)
def test_boxblur_one_handle_many_threads(fmt, args):
    clips = _clips(fmt, 322, 182)
    want = [oa.boxblur(c, **args)["planes"] for c in clips]
    node = vz.core.clip_from_frames(fmt, [c["planes"] for c in clips]).vszip.BoxBlur(**args)
    _hammer(node, lambda n, fr: assert_same_planes(fr.planes, want[n], f"BoxBlur {fmt} {args} frame {n}"))


def test_bilateral_one_handle_many_threads():
    fmt, args = "YUV420P16", dict(sigmaS=2, sigmaR=2)
    clips = _clips(fmt, 322, 182)
    want = [oa.bilateral(c, **args)["planes"] for c in clips]
    node = vz.core.clip_from_frames(fmt, [c["planes"] for c in clips]).vszip.Bilateral(**args)

    def check(n, fr):
        for g, w in zip(fr.planes, want[n]):
            assert np.abs(g.astype(np.int64) - w.astype(np.int64)).max() <= 1, f"Bilateral frame {n}"
    _hammer(node, check)
    # exact-LUT mode (default sigmaR) and PBFIC are bit-exact: any cross-request mix-up of tables or scratch shows up as a wrong bit
    for args in (dict(sigmaS=1.5, sigmaR=0.02), dict(sigmaS=3, sigmaR=0.1, algorithm=1)):
        want = [oa.bilateral(c, **args)["planes"] for c in clips]
        node = vz.core.clip_from_frames(fmt, [c["planes"] for c in clips]).vszip.Bilateral(**args)
        _hammer(node, lambda n, fr: assert_same_planes(fr.planes, want[n], f"Bilateral {args} frame {n}"))


@pytest.mark.parametrize("fmt", ["GRAY16", "GRAYS", "YUV420P8"])
def test_plane_stats_one_handle_many_threads(fmt):
    clips = _clips(fmt, 640, 360)
    planes = None if fmt.startswith("GRAY") else [0, 1, 2]
    excl = [0, 1] if fmt == "GRAYS" else [0, 200]
    src = vz.core.clip_from_frames(fmt, [c["planes"] for c in clips])
    mm_args = dict(minthr=0.1, maxthr=0.05, planes=planes)
    want_mm = [oa.planeminmax(c, **mm_args) for c in clips]
    _hammer(src.vszip.PlaneMinMax(**mm_args), lambda n, fr: _same_props(fr.props, want_mm[n], f"PlaneMinMax {fmt} frame {n}"))
    want_nt = [oa.planeminmax(c, planes=planes) for c in clips]
    _hammer(src.vszip.PlaneMinMax(planes=planes), lambda n, fr: _same_props(fr.props, want_nt[n], f"PlaneMinMax (no thr) {fmt} frame {n}"))
    want_av = [oa.planeaverage(c, excl, planes=planes) for c in clips]
    _hammer(src.vszip.PlaneAverage(exclude=excl, planes=planes), lambda n, fr: _same_props(fr.props, want_av[n], f"PlaneAverage {fmt} frame {n}", rel=1e-12))


def _same_props(got, want, what, rel=0.0):
    for k, v in want.items():
        g = got[k]
        if rel and isinstance(v, float):
            assert g == pytest.approx(v, rel=rel), f"{what}: {k} = {g}, want {v}"
        elif rel and isinstance(v, list):
            assert g == pytest.approx(v, rel=rel), f"{what}: {k} = {g}, want {v}"
        else:
            assert g == v, f"{what}: {k} = {g}, want {v}"


def test_pointwise_and_fused_chain_many_threads():
    fmt = "YUV420P16"
    clips = _clips(fmt, 322, 182)
    src = vz.core.clip_from_frames(fmt, [c["planes"] for c in clips])
    want = [oa.limiter(c, tv_range=True)["planes"] for c in clips]
    _hammer(src.vszip.Limiter(tv_range=True), lambda n, fr: assert_same_planes(fr.planes, want[n], f"Limiter frame {n}"))
    # BoxBlur -> LimitFilter(src) -> PlaneMinMax as ONE fused chain handle (vszip_chain_get_frame) from 16 threads
    node = src.vszip.BoxBlur(hradius=2, vradius=2).vszip.LimitFilter(src, dark_thr=8, bright_thr=4, elast=3).vszip.PlaneMinMax(minthr=0.05, maxthr=0.05)
    wantp = [oa.limitfilter(oa.boxblur(c, hradius=2, vradius=2), c, None, dark_thr=8, bright_thr=4, elast=3) for c in clips]
    wantm = [oa.planeminmax(w, minthr=0.05, maxthr=0.05) for w in wantp]
    vz.core.fuse_chains = True

    def check(n, fr):
        assert_same_planes(fr.planes, wantp[n]["planes"], f"fused chain frame {n}")
        _same_props(fr.props, wantm[n], f"fused chain frame {n}")
    _hammer(node, check)
    assert getattr(node, "_chain", None) is not None
