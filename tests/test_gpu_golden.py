"""The CUDA path (through the C ABI) against the committed golden vectors of tests/golden/ — frozen from the cv2-driven
pipeline / cv2.BFMatcher by tools/make_golden.py, i.e. from the real OpenCV kernels, not from the oracle."""
import os
import sys
import zlib

import numpy as np
import pytest

from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
import make_golden  # noqa: E402

GOLD = os.path.join(HERE, "golden")


@pytest.mark.parametrize("name", sorted(make_golden.EXTRACT_CASES))
def test_extractor_matches_golden(gpu, name):
    kind, h, w, seed, nf, sf, nl, it, mt, lap = make_golden.EXTRACT_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    img = make_golden.make_image(kind, h, w, seed)
    assert zlib.crc32(img.tobytes()) == int(g["crc"])
    ex = ORBextractor(nf, sf, nl, it, mt)
    mono, kps, desc = ex(img, lap)
    assert mono == int(g["mono"]) and len(kps) == len(g["kps"])
    for fld in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(kps[fld].view(np.uint32), g["kps"][fld].view(np.uint32)), "%s: %s" % (name, fld)
    assert np.array_equal(desc, g["desc"])
    if len(kps):
        for l in range(nl):
            assert zlib.crc32(ex.debug_level(l).tobytes()) == int(g["level_crc"][l]), "%s: pyramid level %d" % (name, l)


@pytest.mark.parametrize("name,nq,nt,proto", [("knn2_1k", 1000, 1000, 0), ("knn2_ties", 700, 900, 32)])
def test_knn2_matches_golden(gpu, name, nq, nt, proto):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    q, t = synth.descriptors(nq, 5, proto), synth.descriptors(nt, 6, proto)
    assert zlib.crc32(q.tobytes() + t.tobytes()) == int(g["crc"])
    for got, key in zip(ORBmatcher().knnMatch2(q, t), ("idx1", "d1", "idx2", "d2")):
        assert np.array_equal(got, g[key]), key
