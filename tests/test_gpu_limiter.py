"""Limiter on the GPU against the CPU oracle: bit-exact on every sample type."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes, from_frame, noise_clip, to_node
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "limiter.json").read_text())


def run(clip, **args):
    return from_frame(clip["format"], to_node(clip).vszip.Limiter(**args).get_frame(0))


@pytest.mark.parametrize("key", sorted(k for k in GOLD if k.split("|")[0] in fx.FORMATS))
def test_golden_cases(key):
    fmt, geo, args, _ = oa.parse_case_id(key)
    clip = fx.make_clip(fmt, geo)
    got = run(clip, **args)
    assert_same_planes(got["planes"], oa.limiter(clip, **args)["planes"], key)
    stats = oa.golden_stats(got)
    for p, e in GOLD[key].items():
        assert stats[p]["min"] == e["min"] and stats[p]["max"] == e["max"]
        assert stats[p]["avg"] == pytest.approx(e["avg"], rel=1e-6)  # the reference suite's own tolerance


@pytest.mark.parametrize("fmt", ["GRAY8", "GRAY10", "GRAY16", "GRAY32", "GRAYH", "GRAYS", "YUV420P8", "YUV420P16", "YUV444PS", "RGB24", "RGBS"])
def test_noise(fmt):
    base = "GRAY16" if fmt == "GRAY10" else fmt
    clip = noise_clip(base, 517, 243, seed=61)   # odd width: vector body + scalar row tails
    if fmt == "GRAY10":
        clip = {"format": "GRAY10", "planes": [clip["planes"][0] >> 6]}
    fam, st, bits, ssw, ssh = fx.FORMATS[fmt]
    n = len(clip["planes"])
    peak = (1 << bits) - 1
    if st == "i":
        cases = [dict(min=[peak // 5] * n, max=[peak - peak // 4] * n), dict(min=[0.9] * n, max=[peak] * n)]
    else:
        cases = [dict(min=[0.2] + [-0.3] * (n - 1), max=[0.75] + [0.31] * (n - 1))]
        clip["planes"][0][3, 5] = np.nan   # @max/@min return the other operand for a NaN sample
    cases += [dict(), dict(tv_range=True), dict(mask=True), dict(tv_range=True, mask=True)]
    if n == 3:
        cases += [dict(tv_range=True, planes=[1]), dict(planes=[0, 2])]
    for args in cases:
        got = run(clip, **args)
        assert_same_planes(got["planes"], oa.limiter(clip, **args)["planes"], f"{fmt} {args}")


def test_device_batch_and_chain():
    fmt, w, h, n = "YUV420P16", 640, 360, 3
    a, d = vz.DeviceClip(fmt, w, h, n), vz.DeviceClip(fmt, w, h, n)
    a.fill_noise(seed=5)
    vz.LimiterFilter(a.info(), tv_range=True).run_device(a, d)
    for i in range(n):
        src = {"format": fmt, "planes": a.download(i)}
        assert_same_planes(d.download(i), oa.limiter(src, tv_range=True)["planes"], f"frame {i}")
    clip = noise_clip(fmt, 322, 182, seed=9)
    node = to_node(clip).vszip.BoxBlur(hradius=2, vradius=2).vszip.Limiter(tv_range=True).vszip.PlaneMinMax(minthr=0.01, maxthr=0.01)
    out = node.get_frame(0)                       # fused: one upload, one download
    assert getattr(node, "_chain", None) is not None
    want = oa.limiter(oa.boxblur(clip, hradius=2, vradius=2), tv_range=True)
    assert_same_planes(out.planes, want["planes"], "BoxBlur -> Limiter chain")
    assert out.props["psmMin"] >= 4096 and out.props["psmMax"] <= 60160
