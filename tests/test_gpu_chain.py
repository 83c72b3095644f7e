"""BASELINE config 5: BoxBlur -> Bilateral -> PlaneMinMax chained on device-resident YUV444PS frames
(no PCIe round trips between the filters), against the same chain run through the CPU oracle."""
import numpy as np
import pytest

import oracle_api as oa
import vapoursynth_zip_b200 as vz
from helpers import assert_same_planes

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
