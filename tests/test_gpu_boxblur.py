"""BoxBlur parity on the GPU: the CUDA path, called through the C ABI, against the CPU oracle on
the same inputs.  The bar is bit-exact for every sample type (the kernels execute the reference's
own per-line op sequence), so floats are compared by bit pattern.  Mirrors the structure of the
reference's tests/test_boxblur.py (golden cases, pass composition, h/v composition, planes, stride)."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, from_frame, noise_clip, to_node
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "boxblur.json").read_text())


def run(clip, **args):
    return from_frame(clip["format"], to_node(clip).vszip.BoxBlur(**args).get_frame(0))


# ---- the reference's golden cases, on the GPU: recorded PlaneStats must be reproduced and every
# ---- pixel must equal the oracle's
@pytest.mark.parametrize("key", sorted(GOLD))
def test_golden_cases(key):
    fmt, geo, args, _ = oa.parse_case_id(key)
    clip = fx.make_clip(fmt, geo)
    got = run(clip, **args)
    want = oa.boxblur(clip, **args)
    assert_same_planes(got["planes"], want["planes"], key)
    stats = oa.golden_stats(got)
    for p, e in GOLD[key].items():
        assert stats[p]["min"] == e["min"] and stats[p]["max"] == e["max"]
        assert stats[p]["avg"] == pytest.approx(e["avg"], rel=1e-12)


CASES = [
    # comptime path (hradius == vradius <= 22, one pass)
    dict(hradius=1, vradius=1), dict(hradius=13, vradius=13), dict(hradius=22, vradius=22),
    # runtime path: radius > 22, asymmetric, multi-pass, single axis, > 5 passes (two fused launches)
    dict(hradius=23, vradius=23), dict(hradius=4, vradius=9), dict(hradius=9, vradius=4),
    dict(hradius=13, hpasses=5, vradius=13, vpasses=5),
    dict(hradius=5, hpasses=2, vradius=5, vpasses=1), dict(hradius=5, hpasses=1, vradius=5, vpasses=2),
    dict(hradius=7, vradius=0, vpasses=0), dict(hradius=0, hpasses=0, vradius=7),
    dict(hradius=3, hpasses=6, vradius=2, vpasses=8),
    dict(hradius=30, vradius=33, hpasses=1, vpasses=3),
]


@pytest.mark.parametrize("fmt", ["GRAY8", "GRAY16", "GRAYH", "GRAYS"])
@pytest.mark.parametrize("args", CASES, ids=lambda a: ",".join(f"{k}={v}" for k, v in a.items()))
def test_noise_bit_exact(fmt, args):
    clip = noise_clip(fmt, 331, 203, seed=7)  # odd sizes: partial column groups, partial row blocks, ragged tiles
    assert_same_planes(run(clip, **args)["planes"], oa.boxblur(clip, **args)["planes"], f"{fmt} {args}")


@pytest.mark.parametrize("fmt", ["YUV420P8", "YUV420P10", "YUV420P16", "YUV444PS", "RGBH"])
def test_multi_plane_formats(fmt):
    if fmt == "YUV420P10":
        clip = noise_clip("YUV420P16", 200, 120, seed=3)
        clip = {"format": "YUV420P10", "planes": [(p >> 6) for p in clip["planes"]]}
    else:
        clip = noise_clip(fmt, 200, 120, seed=3)
    for args in (dict(hradius=3, vradius=3), dict(hradius=6, vradius=3, hpasses=2, vpasses=2)):
        assert_same_planes(run(clip, **args)["planes"], oa.boxblur(clip, **args)["planes"], f"{fmt} {args}")


def test_tiny_and_minimum_sizes():
    # 2*radius < dim is the only constraint (src/vapoursynth/boxblur.zig:158-179): smallest legal planes
    for w, h, r in ((3, 3, 1), (5, 3, 1), (13, 7, 2), (27, 27, 13), (47, 47, 23)):
        for fmt in ("GRAY8", "GRAY16", "GRAYS"):
            clip = noise_clip(fmt, w, h, seed=w * h)
            for args in (dict(hradius=r, vradius=r), dict(hradius=r, vradius=r, hpasses=3, vpasses=2)):
                assert_same_planes(run(clip, **args)["planes"], oa.boxblur(clip, **args)["planes"], f"{fmt} {w}x{h} {args}")


@pytest.mark.parametrize("fmt", ["GRAY8", "GRAY16", "GRAYS"])
def test_very_large_radius_uses_global_fallback(fmt):
    """Radii whose 2r+1-slot delay ring cannot fit in shared memory still work (and stay bit-exact)."""
    clip = noise_clip(fmt, 1900, 1000, seed=4)
    for args in (dict(hradius=900, hpasses=2, vradius=480, vpasses=1), dict(hradius=0, hpasses=0, vradius=499, vpasses=2),
                 dict(hradius=940, vradius=0, vpasses=0)):
        assert_same_planes(run(clip, **args)["planes"], oa.boxblur(clip, **args)["planes"], f"{fmt} {args}")


@pytest.mark.parametrize("fmt", ["GRAY8", "GRAY16", "GRAYS"])
def test_pass_composition(fmt):
    """Two passes in one call == two chained single-pass calls, bit-identical (reference test_pass_composition)."""
    src = to_node(fx.make_clip(fmt))
    once = src.vszip.BoxBlur(hradius=7, hpasses=2, vradius=0, vpasses=0).get_frame(0)
    single = dict(hradius=7, hpasses=1, vradius=0, vpasses=0)
    chained = src.vszip.BoxBlur(**single).vszip.BoxBlur(**single).get_frame(0)
    assert_same_planes(once.planes, chained.planes)


def test_h_and_v_compose():
    src = to_node(fx.make_clip("GRAY16"))
    both = src.vszip.BoxBlur(hradius=4, vradius=9).get_frame(0)
    split = src.vszip.BoxBlur(hradius=4, vradius=0, vpasses=0).vszip.BoxBlur(hradius=0, hpasses=0, vradius=9).get_frame(0)
    assert_same_planes(both.planes, split.planes)


def test_planes_subset_shares_untouched_planes():
    clip = fx.make_clip("YUV420P16")
    src = to_node(clip)
    out = src.vszip.BoxBlur(planes=[0], hradius=5, vradius=5).get_frame(0)
    assert out.planes[1] is clip["planes"][1] and out.planes[2] is clip["planes"][2]  # shared, not copied
    want = oa.boxblur(clip, planes=[0], hradius=5, vradius=5)
    assert_same_planes(out.planes, want["planes"])


@pytest.mark.parametrize("radius", [10, 30])
def test_stride_handling(radius):
    """Odd width + offset plane pointer + padded stride give the same pixels as a compact copy."""
    full = fx.make_clip("GRAY16")["planes"][0]
    view = full[:, 27:]                      # non-contiguous rows: stride != width
    a = vz.core.clip_from_frames("GRAY16", [[view]]) if view.strides[1] == view.itemsize else None
    b = vz.core.clip_from_frames("GRAY16", [[np.ascontiguousarray(view)]])
    fa = a.vszip.BoxBlur(hradius=radius, vradius=radius).get_frame(0)
    fb = b.vszip.BoxBlur(hradius=radius, vradius=radius).get_frame(0)
    assert_same_planes(fa.planes, fb.planes)


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (1920x1080 YUV420P16, 13/5/13/5): oracle equality on one frame,
    plus the size-independent identities (5 passes == 2 + 3 chained; constant frames are fixed points)."""
    clip = noise_clip("YUV420P16", 1920, 1080, seed=11)
    args = dict(hradius=13, hpasses=5, vradius=13, vpasses=5)
    got = run(clip, **args)
    assert_same_planes(got["planes"], oa.boxblur(clip, **args)["planes"], "config 2")
    n = to_node(clip)
    a = n.vszip.BoxBlur(hradius=13, hpasses=2, vradius=0, vpasses=0).vszip.BoxBlur(hradius=13, hpasses=3, vradius=0, vpasses=0)
    b = n.vszip.BoxBlur(hradius=13, hpasses=5, vradius=0, vpasses=0)
    assert_same_planes(a.get_frame(0).planes, b.get_frame(0).planes)
    const = vz.core.BlankClip("YUV420P16", 1920, 1080, color=[6777, 32768, 1])
    out = const.vszip.BoxBlur(**args).get_frame(0)
    for p, c in zip(out.planes, (6777, 32768, 1)):
        assert int(p.min()) == int(p.max()) == c


def test_full_size_config1():
    clip = noise_clip("YUV420P16", 1920, 1080, seed=12)
    args = dict(hradius=13, hpasses=1, vradius=13, vpasses=1)
    assert_same_planes(run(clip, **args)["planes"], oa.boxblur(clip, **args)["planes"], "config 1")


def test_device_batch_matches_get_frame():
    """The batched device-resident entry point (what bench.py times) produces the same frames."""
    fmt, w, h, n = "YUV420P16", 256, 144, 5
    src = vz.DeviceClip(fmt, w, h, n)
    dst = vz.DeviceClip(fmt, w, h, n)
    src.fill_noise(seed=99)
    flt = vz.BoxBlurFilter(src.info(), hradius=13, hpasses=5, vradius=13, vpasses=5)
    flt.run_device(src, dst)
    vz.core.sync()
    for i in range(n):
        planes = src.download(i)
        want = oa.boxblur({"format": fmt, "planes": planes}, hradius=13, hpasses=5, vradius=13, vpasses=5)
        assert_same_planes(dst.download(i), want["planes"], f"frame {i}")
    # noise frames differ from each other and use the full range
    a, b = src.download(0)[0], src.download(1)[0]
    assert (a != b).mean() > 0.99 and a.max() > 60000 and a.min() < 5000


# --------------------------------------------------------------------------- comptime float path: streaming-accumulator kernels (boxblur_ctf.cu)
@pytest.mark.parametrize("fmt", ["GRAYS", "GRAYH"])
@pytest.mark.parametrize("r", list(range(1, 23)))
def test_comptime_float_every_radius(fmt, r):
    """Every comptime radius, f32 and f16, on an odd-sized plane (ragged tiles, a partial row block, an odd f16 column pair)
    and on the smallest planes the filter accepts (2r+1 samples per line: interior of one sample); bit-exact incl. the mirrored edge windows."""
    for (w, h) in ((331, 203), (2 * r + 1, 2 * r + 1), (2 * r + 2, 2 * r + 5), (2 * r + 1, 150), (140, 2 * r + 1)):
        clip = noise_clip(fmt, w, h, seed=100 + r)
        assert_same_planes(run(clip, hradius=r, vradius=r)["planes"], oa.boxblur(clip, hradius=r, vradius=r)["planes"], f"{fmt} r={r} {w}x{h}")


@pytest.mark.parametrize(("fmt", "w", "h", "r"), [("GRAYS", 3840, 2160, 13), ("GRAYH", 1920, 1080, 22), ("YUV444PS", 1283, 2047, 5), ("GRAYS", 4099, 517, 1)])
def test_comptime_float_long_lines_are_cut_into_pieces(fmt, w, h, r):
    """Single-frame calls on big planes cut every line into several pieces (each with its own warm-up): piece boundaries,
    the last ragged piece and the signed zeros / denormals of real data must not show."""
    clip = noise_clip(fmt, w, h, seed=31)
    p0 = clip["planes"][0]
    p0[::7, ::5] = 0                       # exact zeros and negative zeros: 0 + (-0) = +0 must be reproduced tap for tap
    p0[3::11, 1::9] = -0.0
    p0[5::13, 2::17] = np.finfo(p0.dtype).tiny / 4   # denormals
    assert_same_planes(run(clip, hradius=r, vradius=r)["planes"], oa.boxblur(clip, hradius=r, vradius=r)["planes"], f"{fmt} {w}x{h} r={r}")


def test_comptime_float_batch_matches_single_frames():
    """Batched launches choose a different cut of the lines than single-frame calls: same bits either way."""
    fmt, w, h, n = "YUV444PS", 640, 360, 40
    src, dst = vz.DeviceClip(fmt, w, h, n), vz.DeviceClip(fmt, w, h, n)
    src.fill_noise(seed=9)
    vz.BoxBlurFilter(src.info(), hradius=13, vradius=13).run_device(src, dst)
    vz.core.sync()
    for i in (0, 17, 39):
        planes = src.download(i)
        want = oa.boxblur({"format": fmt, "planes": planes}, hradius=13, vradius=13)["planes"]
        assert_same_planes(dst.download(i), want, f"frame {i}")


@pytest.mark.parametrize("fmt", ["GRAYS", "GRAYH"])
def test_comptime_float_small_case_for_the_sanitizer(fmt):
    """Two radii on a small plane: the case scripts/sanitize.sh runs under memcheck / racecheck / synccheck (TMA boxes past the
    plane's right and bottom edge, the store warp's partial first and last lines, several pieces per line)."""
    for r in (3, 13):
        clip = noise_clip(fmt, 403, 231, seed=r)
        assert_same_planes(run(clip, hradius=r, vradius=r)["planes"], oa.boxblur(clip, hradius=r, vradius=r)["planes"], f"{fmt} r={r}")


# --------------------------------------------------------------------------- 8-bit clips on the segment kernels (runtime path)
@pytest.mark.parametrize(("fmt", "w", "h"), [("GRAY8", 1920, 1080), ("YUV420P8", 1920, 1080), ("GRAY8", 331, 203), ("YUV420P8", 642, 362),
                                            ("GRAY8", 1280, 720), ("GRAY8", 64, 90), ("GRAY8", 61, 33), ("YUV444P8", 960, 540)])
def test_runtime_8bit_on_segment_kernels(fmt, w, h):
    """8-bit clips run the 16-bit segment arithmetic behind a widen/narrow step: whole and ragged rows, heights that are and are
    not multiples of 90 (TMA tile kernel / plain loads), several radii and pass counts, H-only and V-only."""
    clip = noise_clip(fmt, w, h, seed=w + h)
    for args in (dict(hradius=13, hpasses=5, vradius=13, vpasses=5), dict(hradius=3, hpasses=2, vradius=0, vpasses=0),
                 dict(hradius=0, hpasses=0, vradius=7, vpasses=3), dict(hradius=22, hpasses=1, vradius=1, vpasses=4)):
        if 2 * max(args["hradius"], 1) >= (w >> (1 if fmt.startswith("YUV420") else 0)) or 2 * max(args["vradius"], 1) >= (h >> (1 if fmt.startswith("YUV420") else 0)):
            continue
        assert_same_planes(run(clip, **args)["planes"], oa.boxblur(clip, **args)["planes"], f"{fmt} {w}x{h} {args}")


# --------------------------------------------------------------------------- 8-bit clips on the fused comptime kernel
@pytest.mark.parametrize(("fmt", "w", "h"), [("GRAY8", 1920, 1080), ("YUV420P8", 1920, 1080), ("GRAY8", 331, 203), ("YUV420P8", 642, 362),
                                            ("GRAY8", 2048, 270), ("GRAY8", 64, 90), ("GRAY8", 61, 33), ("YUV444P8", 960, 540), ("GRAY8", 120, 77)])
def test_comptime_8bit_on_fused_kernel(fmt, w, h):
    """hradius == vradius, one pass each (boxblur_comptime.zig) on 8-bit clips: byte rows through the TMA ring, widened on the way
    into the 16-bit column sums and narrowed on the way out; widths that are not multiples of 8 or 16, both column splits (4 and 8
    columns per thread), the defaults (radius 1) and the largest comptime radius."""
    clip = noise_clip(fmt, w, h, seed=3 * w + h)
    sub = 1 if fmt.startswith("YUV420") else 0
    for r in (1, 2, 5, 13, 22):
        if 2 * r >= (w >> sub) or 2 * r >= (h >> sub):
            continue
        assert_same_planes(run(clip, hradius=r, vradius=r)["planes"], oa.boxblur(clip, hradius=r, vradius=r)["planes"], f"{fmt} {w}x{h} r={r}")


@pytest.mark.parametrize("w", [959, 960, 961, 1000, 1024, 1025])
@pytest.mark.parametrize("fmt", ["GRAY16", "GRAY8"])
def test_comptime_widths_around_the_cta_shape_switch(fmt, w):
    """The fused comptime kernel runs 4-warp CTAs up to 960 columns and 8-warp CTAs above (both with 8 columns per thread); 961..1024
    still fits 4 warps' columns but not their rows-per-group rule."""
    clip = noise_clip(fmt, w, 75, seed=w)
    for r in (2, 13):
        assert_same_planes(run(clip, hradius=r, vradius=r)["planes"], oa.boxblur(clip, hradius=r, vradius=r)["planes"], f"{fmt} w={w} r={r}")
