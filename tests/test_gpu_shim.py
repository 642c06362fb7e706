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


def test_threading_contract_two_extractor_threads_and_three_matcher_threads(gpu):
    """SURVEY.md §8(b) threading: the reference runs its left / right extractor objects on two std::threads per frame
    (src/Frame.cc:200-203) and ORBmatcher temporaries on three threads. tests/thread_check.cpp does both on the CUDA
    library (per-thread matcher contexts as in shim/orbx_thread_matcher.h) and compares every output with the oracle's
    stored words (tools/ubench/data/hotpath_case.bin, written by tools/ubench/make_hotpath_case.py)."""
    root = os.path.dirname(HERE)
    case = os.path.join(root, "tools", "ubench", "data", "hotpath_case.bin")
    exe = os.path.join(HERE, "thread_check")
    if not os.path.exists(case) or not os.path.exists(exe):
        pytest.skip("thread_check / its case file are built in the build container (make -C tests; make -C tools/ubench)")
    r = subprocess.run([exe, case, "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "two extractor objects on two threads" in r.stdout and "FAILED" not in r.stdout
