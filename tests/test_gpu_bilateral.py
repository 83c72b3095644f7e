"""Bilateral (algorithm 2 and algorithm 1 / PBFIC) parity on the GPU against the CPU oracle.

Bars (BASELINE.json north_star): integer outputs within 1 LSB with the exact-match fraction reported,
float outputs within 1e-5 relative.  Where the range weights come from the reference's own LUT
(shared-memory LUT: 8..12-bit clips, or any clip whose LUT is short because sigmaR is small; or the
global-LUT mode) the kernel is bit-exact and the tests assert that instead."""
import json
import os
from pathlib import Path

import numpy as np
import pytest

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, from_frame, noise_clip, to_node
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "bilateral.json").read_text())


def run(clip, ref=None, **args):
    node = to_node(clip).vszip.Bilateral(ref=to_node(ref) if ref is not None else None, **args)
    return from_frame(clip["format"], node.get_frame(0)), node.filter.info()


def f16_ordinal(a):
    """f16 bit patterns mapped to integers that are monotonic in the value (so a difference counts ulps)."""
    u = np.ascontiguousarray(a, dtype=np.float16).view(np.uint16).astype(np.int32)
    return np.where(u & 0x8000, -(u & 0x7FFF), u)


def compare(got, want, info, what):
    """exact where the weights are exact, else <= 1 LSB (ints) / 1e-5 relative (floats)."""
    fam, st, bits, ssw, ssh = fx.FORMATS[got["format"]]
    report = []
    for i, (g, w) in enumerate(zip(got["planes"], want["planes"])):
        if not info.process[i] or info.exact_lut[i]:
            assert_same_planes([g], [w], f"{what} plane {i} (exact weights)")
            report.append(1.0)
            continue
        if st == "i":
            d = np.abs(g.astype(np.int64) - w.astype(np.int64))
            assert d.max() <= 1, f"{what} plane {i}: max |diff| = {d.max()} LSB"
            report.append(float((d == 0).mean()))
        else:
            g64, w64 = g.astype(np.float64), w.astype(np.float64)
            if bits == 32:
                # 1e-5 relative (north_star); the absolute floor of 1e-6 of full scale is for samples near zero
                # (float chroma is centred on 0, where a relative bound means nothing)
                tol = 1e-5 * np.abs(w64) + 1e-6
                assert (np.abs(g64 - w64) <= tol).all(), f"{what} plane {i}: max err {np.abs(g64 - w64).max()}"
            else:
                # f16 output: 1e-5 relative is far below half an f16 ulp (4.9e-4 relative), so the bound is "the f32 result
                # rounds to the same or the neighbouring f16": at most 1 ulp, exact-match fraction reported
                ulps = np.abs(f16_ordinal(g) - f16_ordinal(w))
                assert ulps.max() <= 1, f"{what} plane {i}: {int(ulps.max())} f16 ulps"
            report.append(float((g64 == w64).mean()))
    return report


@pytest.mark.parametrize("key", sorted(k for k in GOLD if "|ref" not in k))
def test_golden_cases(key):
    fmt, geo, args, _ = oa.parse_case_id(key)
    clip = fx.make_clip(fmt, geo)
    got, info = run(clip, **args)
    want = oa.bilateral(clip, **args)
    frac = compare(got, want, info, key)
    print(f"\n[bilateral] {key}: exact-match fraction per plane = {frac}")
    stats = oa.golden_stats(got)
    for p, e in GOLD[key].items():
        assert stats[p]["avg"] == pytest.approx(e["avg"], rel=1e-6)  # the reference suite's own tolerance


@pytest.mark.parametrize("fmt", ["GRAY8", "GRAY16", "GRAYH", "GRAYS", "YUV420P16", "YUV444PS"])
@pytest.mark.parametrize("args", [dict(sigmaS=2, sigmaR=2), dict(), dict(sigmaS=0.8, sigmaR=0.02), dict(sigmaS=5, sigmaR=0.02),
                                  dict(sigmaS=5, sigmaR=2)], ids=str)
def test_noise(fmt, args):
    clip = noise_clip(fmt, 203, 131, seed=5)
    got, info = run(clip, **args)
    frac = compare(got, oa.bilateral(clip, **args), info, f"{fmt} {args}")
    print(f"\n[bilateral] noise {fmt} {args}: exact-match fraction per plane = {frac}")


def test_small_sigma_r_is_bit_exact_16bit():
    """Default sigmaR=0.02 on 16 bit: the LUT has 10486 live entries and sits in shared memory."""
    clip = noise_clip("GRAY16", 320, 200, seed=6)
    clip["planes"][0] = (clip["planes"][0] >> 3) + 20000  # keep |a-b| inside the live LUT range too
    got, info = run(clip, sigmaS=3, sigmaR=0.02, algorithm=2)
    assert info.exact_lut[0] == 1
    assert_same_planes(got["planes"], oa.bilateral(clip, sigmaS=3, sigmaR=0.02, algorithm=2)["planes"])


def test_joint_ref():
    clip = fx.make_clip("GRAY16")
    ref = oa.boxblur(clip, hradius=5, vradius=5)
    got, info = run(clip, ref=ref, sigmaS=2, sigmaR=0.05)
    want = oa.bilateral(clip, ref=ref, sigmaS=2, sigmaR=0.05)
    compare(got, want, info, "joint")


def test_sigma_zero_is_passthrough():
    clip = fx.make_clip("GRAY16")
    src = to_node(clip)
    for kw in (dict(sigmaS=0), dict(sigmaR=0)):
        out = src.vszip.Bilateral(**kw).get_frame(0)
        assert out.planes[0] is clip["planes"][0]


def test_planes_and_stride():
    clip = fx.make_clip("YUV420P16")
    out = to_node(clip).vszip.Bilateral(sigmaS=2, sigmaR=2, planes=[0]).get_frame(0)
    assert out.planes[1] is clip["planes"][1]
    full = fx.make_clip("GRAY16")["planes"][0]
    view = full[:, 27:]
    a = vz.core.clip_from_frames("GRAY16", [[view]]).vszip.Bilateral(sigmaS=2, sigmaR=2).get_frame(0)
    b = vz.core.clip_from_frames("GRAY16", [[np.ascontiguousarray(view)]]).vszip.Bilateral(sigmaS=2, sigmaR=2).get_frame(0)
    assert_same_planes(a.planes, b.planes)


def test_full_size_config3():
    """BASELINE config 3: 1920x1080 YUV420P16, sigmaS=2 sigmaR=2 planes=[0,1,2]."""
    clip = noise_clip("YUV420P16", 1920, 1080, seed=21)
    got, info = run(clip, sigmaS=2, sigmaR=2, planes=[0, 1, 2])
    frac = compare(got, oa.bilateral(clip, sigmaS=2, sigmaR=2, planes=[0, 1, 2]), info, "config 3")
    print(f"\n[bilateral] config 3 exact-match fraction per plane = {frac}")
    assert min(frac) > 0.9


def test_device_batch_matches_get_frame():
    fmt, w, h, n = "YUV420P16", 256, 144, 3
    src, dst = vz.DeviceClip(fmt, w, h, n), vz.DeviceClip(fmt, w, h, n)
    src.fill_noise(seed=5)
    flt = vz.BilateralFilter(src.info(), sigmaS=2, sigmaR=2)
    flt.run_device(src, dst)
    vz.core.sync()
    for i in range(n):
        planes = src.download(i)
        node = vz.core.clip_from_frames(fmt, [planes]).vszip.Bilateral(sigmaS=2, sigmaR=2)
        assert_same_planes(dst.download(i), node.get_frame(0).planes, f"frame {i}")


# --------------------------------------------------------------------------- algorithm 1 (PBFIC): bit-exact
@pytest.mark.parametrize("fmt", ["GRAY8", "GRAY10", "GRAY16", "GRAYH", "GRAYS", "YUV420P16", "YUV444PS"])
@pytest.mark.parametrize("args", [dict(sigmaS=3, sigmaR=0.1, PBFICnum=4), dict(sigmaS=8, sigmaR=0.1), dict(sigmaS=1.5, sigmaR=0.02, PBFICnum=7),
                                  dict(sigmaS=[4, 2], sigmaR=[0.2, 0.05], PBFICnum=[3, 9])], ids=str)
def test_pbfic_noise(fmt, args):
    base = "GRAY16" if fmt == "GRAY10" else fmt
    clip = noise_clip(base, 203, 131, seed=15)
    if fmt == "GRAY10":
        clip = {"format": "GRAY10", "planes": [clip["planes"][0] >> 6]}
    got, info = run(clip, algorithm=1, **args)
    assert all(a == 1 for a, p in zip(info.algorithm, info.process) if p)
    assert_same_planes(got["planes"], oa.bilateral(clip, algorithm=1, **args)["planes"], f"PBFIC {fmt} {args}")


def test_pbfic_auto_selected_and_mixed_planes():
    """sigmaS=8, sigmaR=0.1 auto-selects algorithm 1 on luma (bilateral.zig:196); chroma of a 4:2:0 clip gets
    sigmaS/2 and stays on algorithm 2 (computed weights there: <= 1 LSB)."""
    clip = noise_clip("YUV420P16", 322, 178, seed=21)
    got, info = run(clip, sigmaS=8, sigmaR=0.1)
    want = oa.bilateral(clip, sigmaS=8, sigmaR=0.1)
    assert list(info.algorithm)[0] == 1
    compare(got, want, info, "auto PBFIC")


def test_pbfic_joint_and_tiny_planes():
    clip = noise_clip("GRAY16", 97, 45, seed=3)
    ref = noise_clip("GRAY16", 97, 45, seed=4)
    got, _ = run(clip, ref=ref, sigmaS=3, sigmaR=0.1, algorithm=1)
    assert_same_planes(got["planes"], oa.bilateral(clip, ref=ref, sigmaS=3, sigmaR=0.1, algorithm=1)["planes"], "PBFIC joint")
    for (w, h) in ((1, 1), (1, 9), (9, 1), (2, 3), (33, 2), (64, 33)):
        c = noise_clip("GRAYS", w, h, seed=w * 100 + h)
        got, _ = run(c, sigmaS=2, sigmaR=0.3, algorithm=1, PBFICnum=2)
        assert_same_planes(got["planes"], oa.bilateral(c, sigmaS=2, sigmaR=0.3, algorithm=1, PBFICnum=2)["planes"], f"PBFIC {w}x{h}")


def test_pbfic_device_batch_and_full_size():
    fmt, w, h, n = "YUV420P16", 1920, 1080, 3
    a, d = vz.DeviceClip(fmt, w, h, n), vz.DeviceClip(fmt, w, h, n)
    a.fill_noise(seed=11)
    f = vz.BilateralFilter(a.info(), sigmaS=3, sigmaR=0.1, algorithm=1, planes=[0, 1])
    f.run_device(a, d)
    src = {"format": fmt, "planes": a.download(2)}
    want = oa.bilateral(src, sigmaS=3, sigmaR=0.1, algorithm=1, planes=[0, 1])
    got = d.download(2)
    assert_same_planes(got[:2], want["planes"][:2], "PBFIC 1080p frame 2")
