"""CPU-only: the C-ABI library loads and exports exactly what include/vszip_cuda.h declares, and the
product never reaches into oracle/.  No compute call is made here (there is no GPU in this container)."""
import ctypes
import re
import subprocess
from pathlib import Path

import vapoursynth_zip_b200 as vz

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "vszip_cuda.h").read_text()


def declared_symbols():
    # every function declaration in the header: "<ret> vszip_xxx(" at the start of a line
    names = set(re.findall(r"^[A-Za-z_][\w\s\*]*?\b(vszip_\w+)\s*\(", HEADER, flags=re.M))
    return {n for n in names if not n.endswith("_args") and not n.endswith("_props")}


def test_header_symbols_are_exported_and_bound():
    lib = ctypes.CDLL(str(vz.LIB_PATH))
    decl = declared_symbols()
    assert len(decl) >= 28
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in vszip_cuda.h but not exported"
    assert decl == set(vz.ABI), f"python binding and header disagree: {decl ^ set(vz.ABI)}"
    assert vz.load_library().vszip_cuda_abi_version() == 5


def test_signatures_are_plain_c():
    """No torch / CUDA / C++ types in the boundary."""
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)  # declarations only, comments stripped
    for bad in ("torch", "at::", "cudaStream_t", "std::", "__half", "template", "class "):
        assert bad not in code, bad


def test_library_has_no_oracle_dependency():
    out = subprocess.run(["ldd", str(vz.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for src in (ROOT / "vapoursynth_zip_b200").rglob("*"):
        if src.suffix in (".py", ".cu", ".h", ".cpp", ".cuh"):
            text = src.read_text()
            assert "import oracle" not in text and "from oracle" not in text and "vso_" not in text, src


def test_no_gpu_means_loud_failure():
    """Without a CUDA device every compute entry point must fail with a message, never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    clip = vz.core.BlankClip("GRAY16", 64, 64)
    node = clip.vszip.BoxBlur(hradius=2, vradius=2)
    try:
        node.get_frame(0)
    except vz.Error as e:
        assert "no CUDA device" in str(e) or "no CPU fallback" in str(e)
    else:
        raise AssertionError("BoxBlur produced a frame without a GPU")


def test_kernels_are_sm100a_only():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", str(vz.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_integration_source_list_matches_the_makefile():
    """INTEGRATION.md's build.zig fragment must name every .cu the Makefile links, or the maintainer's build would not link."""
    import re
    mk = (ROOT / "Makefile").read_text()
    srcs = re.search(r"^SRCS := (.*)$", mk, re.M).group(1).split()
    doc = (ROOT / "INTEGRATION.md").read_text()
    block = re.search(r"const cuda_srcs = \[_\]\[\]const u8\{(.*?)\};", doc, re.S).group(1)
    assert sorted(re.findall(r'"(\w+)"', block)) == sorted(s[:-3] for s in srcs)
    assert sorted(p.name for p in (ROOT / "vapoursynth_zip_b200" / "csrc").glob("*.cu")) == sorted(srcs)
