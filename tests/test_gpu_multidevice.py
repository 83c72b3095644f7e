"""In-process multi-GPU routing (frame n -> device n mod k): needs >= 2 visible GPUs, skipped otherwise.
Runs in a subprocess because the library is initialised once per process with its device list."""
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

SCRIPT = textwrap.dedent("""
    import sys
    sys.path.insert(0, %r); sys.path.insert(0, %r)
    import numpy as np
    import oracle_api as oa
    import vapoursynth_zip_b200 as vz
    from helpers import assert_same_planes, noise_clip
    k = vz.core.init([0, 1])
    assert k == 2 and vz.core.num_devices == 2
    frames = [noise_clip("YUV420P16", 322, 182, seed=100 + i)["planes"] for i in range(6)]
    clip = vz.core.clip_from_frames("YUV420P16", frames)
    blur = clip.vszip.BoxBlur(hradius=3, hpasses=2, vradius=2, vpasses=2)
    bil = clip.vszip.Bilateral(sigmaS=8, sigmaR=0.1)          # PBFIC luma: per-device LUT upload + scratch
    mm = blur.vszip.PlaneMinMax(minthr=0.1, maxthr=0.1)       # fused chain on either device
    for n in range(6):                                        # even frames -> device 0, odd -> device 1
        src = {"format": "YUV420P16", "planes": frames[n]}
        assert_same_planes(blur.get_frame(n).planes, oa.boxblur(src, hradius=3, hpasses=2, vradius=2, vpasses=2)["planes"], f"blur {n}")
        want = oa.bilateral(src, sigmaS=8, sigmaR=0.1)
        got = bil.get_frame(n).planes
        assert np.array_equal(got[0], want["planes"][0]), n
        p = mm.get_frame(n).props
        w = oa.planeminmax(oa.boxblur(src, hradius=3, hpasses=2, vradius=2, vpasses=2), minthr=0.1, maxthr=0.1)
        assert p["psmMin"] == w["psmMin"] and p["psmMax"] == w["psmMax"], (n, p, w)
    # device-resident clips on device 1
    a, d = vz.DeviceClip("GRAY16", 256, 128, 2, device=1), vz.DeviceClip("GRAY16", 256, 128, 2, device=1)
    a.fill_noise(seed=3)
    vz.BoxBlurFilter(a.info(), hradius=2, vradius=2).run_device(a, d)
    assert_same_planes(d.download(1), oa.boxblur({"format": "GRAY16", "planes": a.download(1)}, hradius=2, vradius=2)["planes"], "dev 1")
    print("OK")
""") % (str(ROOT), str(ROOT / "tests"))


def test_two_devices_in_one_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
