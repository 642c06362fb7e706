"""The header-compatible C++ shim (shim/ORBextractor.{h,cc}) driven like Frame::ExtractORB drives the reference class
(src/Frame.cc:549-560): std::vector<cv::KeyPoint>, cv::Mat descriptors, monoIndex and the mvImagePyramid mirror must
equal the oracle's. The shim is compiled against a minimal mock of the OpenCV types (tests/mock_cv) because the image
has no OpenCV C++ headers; the memory layouts that cross the ABI are OpenCV's."""
import os
import subprocess

import numpy as np
import pytest

from orb_slam3_fast_b200 import synth
from oracle import orbref

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("w,h,nf,lap", [(640, 480, 1000, (0, 1000)), (752, 480, 1200, (0, 0))])
def test_cpp_shim_matches_oracle(gpu, tmp_path, w, h, nf, lap):
    subprocess.check_call(["make", "-C", HERE, "-s", "shim_smoke"])
    img = synth.scene(h, w, seed=21)
    raw, out = tmp_path / "in.raw", tmp_path / "out.bin"
    img.tofile(raw)
    txt = subprocess.check_output([os.path.join(HERE, "shim_smoke"), str(raw), str(w), str(h), str(nf), str(lap[0]),
                                   str(lap[1]), str(out)], text=True)
    buf = open(out, "rb").read()
    mono, n = np.frombuffer(buf, np.int32, 2)
    kps = np.frombuffer(buf, orbref.KP_DTYPE, n, 8)
    off = 8 + 28 * n
    desc = np.frombuffer(buf, np.uint8, 32 * n, off).reshape(n, 32)
    off += 32 * n
    lw, lh = np.frombuffer(buf, np.int32, 2, off)
    bordered = np.frombuffer(buf, np.uint8, (lw + 38) * (lh + 38), off + 8).reshape(lh + 38, lw + 38)
    ex = orbref.Extractor(nf)
    mono_r, kps_r, desc_r = ex(img, lap)
    assert (mono, n) == (mono_r, len(kps_r)), txt
    assert np.array_equal(kps, kps_r) and np.array_equal(desc, desc_r)
    assert np.array_equal(bordered, ex.level_bordered(3))
    assert "sf1=1.20000005" in txt   # GetScaleFactors()[1] = float(1.0 * double(1.2f))
