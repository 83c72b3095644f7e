"""Shared helpers for the parity tests: oracle clips <-> product clips, exact comparison."""
import numpy as np

import vapoursynth_zip_b200 as vz
from oracle import fixtures as fx


def to_node(clip) -> "vz.VideoNode":
    """oracle-style clip dict -> product VideoNode (single frame)."""
    return vz.core.clip_from_frames(clip["format"], [clip["planes"]])


def from_frame(fmt_name, frame) -> dict:
    return {"format": fmt_name, "planes": [np.ascontiguousarray(p) for p in frame.planes]}


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype.kind == "f":
        ua = a.view({2: np.uint16, 4: np.uint32}[a.itemsize])
        ub = b.view({2: np.uint16, 4: np.uint32}[b.itemsize])
        # +0.0 / -0.0 and NaN payloads must match too: compare raw bits
        return bool(np.array_equal(ua, ub))
    return bool(np.array_equal(a, b))


def assert_same_planes(got, want, what=""):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        if bits_equal(g, w):
            continue
        assert g.shape == w.shape and g.dtype == w.dtype, f"{what} plane {i}: {g.shape}/{g.dtype} vs {w.shape}/{w.dtype}"
        raw = {1: np.uint8, 2: np.uint16, 4: np.uint32}[g.itemsize]
        diff = np.argwhere(g.view(raw) != w.view(raw))
        first = tuple(diff[0])
        raise AssertionError(f"{what} plane {i}: {len(diff)} of {g.size} samples differ; first at (y,x)={first}: "
                             f"got {g[first]!r} want {w[first]!r}")


def noise_clip(fmt_name, width, height, seed=0):
    """Host-side random clip in the oracle's clip format (full-range ints, [0,1) / [-0.5,0.5) floats)."""
    fam, st, bits, ssw, ssh = fx.FORMATS[fmt_name]
    rng = np.random.default_rng(seed)
    nplanes = 1 if fam == "GRAY" else 3
    planes = []
    for p in range(nplanes):
        w, h = (width >> ssw, height >> ssh) if p else (width, height)
        if st == "i":
            dt = np.uint8 if bits <= 8 else (np.uint16 if bits <= 16 else np.uint32)
            planes.append(rng.integers(0, 1 << bits, size=(h, w), dtype=np.uint64).astype(dt))
        else:
            v = rng.random((h, w), dtype=np.float32)
            if p and fam == "YUV":
                v = v - np.float32(0.5)
            planes.append(v.astype(np.float16 if bits == 16 else np.float32))
    return {"format": fmt_name, "planes": planes}
