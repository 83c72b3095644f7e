"""The segment BoxBlur kernels' per-thread arithmetic (csrc/boxblur_seg_core.h), replayed lane by lane on the CPU
(tests/sim/boxblur_seg_sim.cpp), against the oracle: bit-exact for every radius the kernels are instantiated for,
including ragged last segments, minimum-size lines and band restarts of the comptime path."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle

ROOT = Path(__file__).resolve().parents[1]
SIM_SRC = ROOT / "tests" / "sim" / "boxblur_seg_sim.cpp"
SIM_LIB = ROOT / "build" / "libsegsim.so"


@pytest.fixture(scope="module")
def sim():
    SIM_LIB.parent.mkdir(exist_ok=True)
    core = ROOT / "vapoursynth_zip_b200" / "csrc" / "boxblur_seg_core.h"
    if not SIM_LIB.exists() or SIM_LIB.stat().st_mtime < max(SIM_SRC.stat().st_mtime, core.stat().st_mtime):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", f"-I{core.parent}", "-o", str(SIM_LIB), str(SIM_SRC)], check=True)
    lib = C.CDLL(str(SIM_LIB))
    lib.seg_sim_u16.argtypes = [C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 5

    def run(what, src, r, passes=1, opt=0):
        src = np.ascontiguousarray(src)
        dst = np.zeros_like(src)
        h, w = src.shape
        assert lib.seg_sim_u16(what, src.ctypes.data, dst.ctypes.data, w, h, r, passes, opt) == 0
        return dst
    return run


def noise(w, h, seed, bits=16):
    return np.random.default_rng(seed).integers(0, 1 << bits, size=(h, w), dtype=np.uint32).astype(np.uint16)


@pytest.mark.parametrize("r", list(range(1, 23)))
def test_h_and_v_passes_match_oracle(sim, r):
    for w, h, passes, opt in ((2 * r + 1, 2 * r + 1, 2, 0), (61, 59, 1, 1), (331, 203, 3, 2), (120, 60, 2, 1), (37, 90, 2, 0), (9, 270, 3, 1), (60, 47, 2, 2), (180, 47, 3, 1)):
        if w <= 2 * r or h <= 2 * r:
            continue
        src = noise(w, h, seed=r * 100 + passes)
        assert np.array_equal(sim(0, src, r, passes, opt), oracle.boxblur_plane(src, r, passes, 0, 0)), ("H", r, w, h, passes)
        assert np.array_equal(sim(1, src, r, passes, opt), oracle.boxblur_plane(src, 0, 0, r, passes)), ("V", r, w, h, passes)


@pytest.mark.parametrize("r", [1, 2, 7, 13, 22])
def test_comptime_path_matches_oracle(sim, r):
    for w, h, band in ((2 * r + 1, 2 * r + 1, 0), (97, 2 * r + 2, 3), (200, 131, 40), (64, 300, 16)):
        src = noise(w, h, seed=r)
        want = oracle.boxblur_plane(src, r, 1, r, 1)
        assert np.array_equal(sim(2, src, r, 1, band), want), ("CT", r, w, h, band)


def test_headline_geometry(sim):
    """config 2's luma row length and the chroma one, five passes; small sample ranges too (10-bit)."""
    for w, bits in ((1920, 16), (960, 16), (1920, 10)):
        src = noise(w, 61, seed=w, bits=bits)
        assert np.array_equal(sim(0, src, 13, 5), oracle.boxblur_plane(src, 13, 5, 0, 0))
    src = noise(70, 1080, seed=5)
    assert np.array_equal(sim(1, src, 13, 5), oracle.boxblur_plane(src, 0, 0, 13, 5))
    src = noise(66, 540, seed=6)
    assert np.array_equal(sim(1, src, 13, 5, 1), oracle.boxblur_plane(src, 0, 0, 13, 5))
