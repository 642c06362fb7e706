"""The oracle's whole extractor vs (a) the committed goldens frozen from the cv2-driven pipeline
(tools/make_golden.py) and (b) that pipeline run live on small cases. Bit-exact on every field.

The goldens are the parity anchor that travels: the `-m gpu` tests compare the CUDA path with the same files."""
import glob
import os
import sys
import zlib

import numpy as np
import pytest

from oracle import orbref

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
import make_golden  # noqa: E402

GOLD = os.path.join(HERE, "golden")


def assert_same_frame(mono, kps, desc, g_mono, g_kps, g_desc, tag):
    assert mono == g_mono, "%s: monoIndex %d vs %d" % (tag, mono, g_mono)
    assert len(kps) == len(g_kps), "%s: N %d vs %d" % (tag, len(kps), len(g_kps))
    for fld in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        a, b = kps[fld], g_kps[fld]
        bad = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0] if a.dtype == np.float32 else np.nonzero(a != b)[0]
        assert len(bad) == 0, "%s: %s differs at rows %s" % (tag, fld, bad[:8])
    bad = np.nonzero((desc != g_desc).any(axis=1))[0]
    assert len(bad) == 0, "%s: %d descriptors differ (rows %s)" % (tag, len(bad), bad[:8])


def test_goldens_exist_for_every_case():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz"))}
    assert set(make_golden.EXTRACT_CASES) <= have and {"knn2_1k", "knn2_ties"} <= have


@pytest.mark.parametrize("name", sorted(make_golden.EXTRACT_CASES))
def test_oracle_extractor_matches_golden(name):
    kind, h, w, seed, nf, sf, nl, it, mt, lap = make_golden.EXTRACT_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    img = make_golden.make_image(kind, h, w, seed)
    assert zlib.crc32(img.tobytes()) == int(g["crc"]), "synthetic generator drifted: regenerate the goldens"
    ex = orbref.Extractor(nf, sf, nl, it, mt)
    mono, kps, desc = ex(img, lap)
    assert_same_frame(mono, kps, desc, int(g["mono"]), g["kps"], g["desc"], name)
    for l in range(nl):
        assert zlib.crc32(ex.level_image(l).tobytes()) == int(g["level_crc"][l]), "%s: pyramid level %d" % (name, l)


@pytest.mark.parametrize("name,nq,nt,proto", [("knn2_1k", 1000, 1000, 0), ("knn2_ties", 700, 900, 32)])
def test_oracle_knn2_matches_golden(name, nq, nt, proto):
    from orb_slam3_fast_b200 import synth
    g = np.load(os.path.join(GOLD, name + ".npz"))
    q, t = synth.descriptors(nq, 5, proto), synth.descriptors(nt, 6, proto)
    assert zlib.crc32(q.tobytes() + t.tobytes()) == int(g["crc"])
    for got, key in zip(orbref.knn2(q, t), ("idx1", "d1", "idx2", "d2")):
        assert np.array_equal(got, g[key]), key


@pytest.mark.parametrize("kind,h,w,seed,nf,lap", [("scene", 241, 241, 9, 300, (0, 0)),
                                                   ("noise_blur", 260, 300, 4, 400, (50, 200)),
                                                   ("uniform_noise", 250, 330, 5, 200, (0, 1000))])
def test_oracle_extractor_matches_live_cv2_pipeline(kind, h, w, seed, nf, lap):
    pytest.importorskip("cv2")
    from oracle import cv2_pipeline
    img = make_golden.make_image(kind, h, w, seed)
    mono, kps, desc = orbref.Extractor(nf)(img, lap)
    g_mono, g_kps, g_desc = cv2_pipeline.Cv2Extractor(nf)(img, lap)
    assert_same_frame(mono, kps, desc, g_mono, g_kps, g_desc, "%s %dx%d" % (kind, w, h))


def test_oracle_edge_cases():
    from orb_slam3_fast_b200 import synth
    ex = orbref.Extractor(1000)
    mono, kps, desc = ex(synth.constant(480, 640), (0, 0))
    assert mono == 0 and len(kps) == 0 and desc.shape == (0, 32)       # descriptors released (:1050-1052)
    mono, kps, desc = ex(np.empty((0, 0), np.uint8))
    assert mono == -1                                                   # empty image (:1021)
    # setup arithmetic (:418-448): quotas of the three benchmark configurations
    assert orbref.Extractor(1000).features_per_level.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert orbref.Extractor(1200).features_per_level.tolist() == [261, 217, 181, 151, 126, 105, 87, 72]
    assert orbref.Extractor(2000).features_per_level.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert orbref.Extractor(1000).umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    ex(synth.scene(480, 640, 0), (0, 0))
    assert [ex.level_dims(l) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231),
                                                     (257, 193), (214, 161), (179, 134)]
