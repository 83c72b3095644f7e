"""BASELINE config 5 at its real size: BoxBlur -> Bilateral -> PlaneMinMax on 3840x2160 YUV444PS, through the batched
`*_device` entry points and through the fused `vszip_chain_get_frame` frame API, against the CPU oracle.

Bars (BASELINE.json north_star): BoxBlur bit-exact (also for f32), Bilateral within 1e-5 relative, PlaneMinMax exact.
The absolute floor of 1e-6 next to the relative bound exists for float chroma only (centred on 0: a sample that the blur
brought to |v| < 0.1 has no meaningful relative error); luma is held to the pure relative bound."""
import numpy as np
import pytest

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, noise_clip, to_node

pytestmark = pytest.mark.gpu
FMT, W, H = "YUV444PS", 3840, 2160


def _check_bilateral(got, want, what):
    for i, (g, w) in enumerate(zip(got, want)):
        g64, w64 = g.astype(np.float64), w.astype(np.float64)
        err = np.abs(g64 - w64)
        tol = 1e-5 * np.abs(w64) + (1e-6 if i else 0.0)
        worst = float((err / np.maximum(np.abs(w64), 1e-30)).max()) if i == 0 else float(err.max())
        assert (err <= tol).all(), f"{what} plane {i}: worst {'relative' if i == 0 else 'absolute'} error {worst:.3g}"
        print(f"\n[config 5] {what} plane {i}: exact {float((g64 == w64).mean()):.4f}, worst {'rel' if i == 0 else 'abs'} err {worst:.3g}")


@pytest.fixture(scope="module")
def expected():
    clip = noise_clip(FMT, W, H, seed=505)
    blur = oa.boxblur(clip, hradius=13, vradius=13)
    bil = oa.bilateral(blur, sigmaS=2, sigmaR=2)
    return clip, blur, bil


def test_config5_device_entry_points_full_size(expected):
    clip, want_blur, want_bil = expected
    n = 2
    a, b, c = (vz.DeviceClip(FMT, W, H, n) for _ in range(3))
    a.upload(0, clip["planes"])
    a.upload(1, [p[::-1].copy() for p in clip["planes"]])      # second frame of the batch: same content upside down
    blur = vz.BoxBlurFilter(a.info(), hradius=13, hpasses=1, vradius=13, vpasses=1)
    bil = vz.BilateralFilter(a.info(), sigmaS=2, sigmaR=2)
    mm = vz.PlaneMinMaxFilter(a.info(), minthr=0.1, maxthr=0.1, planes=[0])
    blur.run_device(a, b)
    bil.run_device(b, c)
    props = mm.run_device(c)
    got_blur = b.download(0)
    assert_same_planes(got_blur, want_blur["planes"], "config 5 BoxBlur 3840x2160 (bit-exact f32)")
    # BoxBlur's SYM/R101 mirroring is symmetric under a vertical flip only up to the summation order, so frame 1 is checked
    # through the reduction below rather than against a flipped oracle frame
    got_bil = c.download(0)
    _check_bilateral(got_bil, want_bil["planes"], "Bilateral (device)")
    for i in range(n):
        frame = c.download(i)
        assert props[i] == oa.planeminmax({"format": FMT, "planes": frame}, minthr=0.1, maxthr=0.1, planes=[0]), f"frame {i}"
    for d in (a, b, c):
        d.free()


def test_config5_fused_chain_full_size(expected):
    clip, want_blur, want_bil = expected
    node = to_node(clip).vszip.BoxBlur(hradius=13, vradius=13).vszip.Bilateral(sigmaS=2, sigmaR=2).vszip.PlaneMinMax(minthr=0.1, maxthr=0.1, planes=[0])
    vz.core.fuse_chains = True
    out = node.get_frame(0)
    assert getattr(node, "_chain", None) is not None, "the chain was not fused"
    _check_bilateral(out.planes, want_bil["planes"], "Bilateral (fused chain)")
    want_mm = oa.planeminmax({"format": FMT, "planes": [np.ascontiguousarray(p) for p in out.planes]}, minthr=0.1, maxthr=0.1, planes=[0])
    assert {k: out.props[k] for k in want_mm} == want_mm
