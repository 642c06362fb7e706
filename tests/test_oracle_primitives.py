"""The oracle's restated third-party arithmetic vs the REAL OpenCV 4.13 kernels (cv2), glibc and libstdc++.

This is what pins the oracle (the reference has no tests of its own and cannot be built here; its hot
path spends most of its arithmetic inside OpenCV calls — src/ORBextractor.cc:78,98,810,820,1075,1122,1129,1137 and
src/Frame.cc:1293). Every comparison is byte / bit exact.
"""
import ctypes

import numpy as np
import pytest

from orb_slam3_fast_b200 import synth
from oracle import orbref

cv2 = pytest.importorskip("cv2")


def _images():
    yield "scene", synth.scene(480, 640, 0)
    yield "noise", synth.uniform_noise(333, 444, 1)
    yield "blur", synth.noise_blur(241, 257, 2)


@pytest.mark.parametrize("scale", [1.2, 1.1, 1.5, 2.0])
def test_resize_linear_is_cv_resize(scale):
    # cv::resize(level l-1 -> sz, INTER_LINEAR), src/ORBextractor.cc:1122
    for name, img in _images():
        h, w = img.shape
        for dw, dh in ((int(round(w / scale)), int(round(h / scale))), (w - 1, h - 3)):
            ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
            got = orbref.resize_linear(img, dw, dh)
            assert np.array_equal(got, ref), "%s %dx%d -> %dx%d: %d px differ" % (name, w, h, dw, dh, (got != ref).sum())


def test_resize_chain_is_the_pyramid():
    img = synth.scene(480, 752, 4)
    ex = orbref.Extractor(1200)
    ex(img, (0, 0))
    prev = img
    for l in range(1, 8):
        w, h = ex.level_dims(l)
        prev = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ex.level_image(l), prev), "level %d" % l


def test_gauss7_is_cv_gaussianblur():
    # GaussianBlur(workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101), src/ORBextractor.cc:1075
    for name, img in _images():
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        got = orbref.gauss7(img)
        assert np.array_equal(got, ref), "%s: %d px differ" % (name, (got != ref).sum())
    tiny = synth.uniform_noise(9, 11, 3)  # reflect-101 dominates
    assert np.array_equal(orbref.gauss7(tiny), cv2.GaussianBlur(tiny, (7, 7), 2, sigmaY=2,
                                                                 borderType=cv2.BORDER_REFLECT_101))


def test_border101_is_copymakeborder():
    # copyMakeBorder(..., EDGE_THRESHOLD x4, BORDER_REFLECT_101), src/ORBextractor.cc:1129-1143
    for name, img in _images():
        ref = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        assert np.array_equal(orbref.border101(img, 19), ref), name


@pytest.mark.parametrize("threshold", [20, 7, 1, 40])
def test_fast9_is_cv_fast(threshold):
    # cv::FAST(cell, keys, threshold, true), src/ORBextractor.cc:810,820 (TYPE_9_16 is the default)
    det = cv2.FastFeatureDetector_create(threshold, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    imgs = [synth.scene(120, 160, 0), synth.uniform_noise(60, 80, 1), synth.noise_blur(44, 42, 2),
            synth.scene(480, 640, 5)[100:144, 200:242].copy(), synth.low_contrast(96, 96, 1, 40)]
    total = 0
    for img in imgs:
        kps = det.detect(img)
        xs, ys, sc = orbref.fast9(img, threshold)
        assert len(kps) == len(xs)
        assert [int(k.pt[0]) for k in kps] == xs.tolist() and [int(k.pt[1]) for k in kps] == ys.tolist()
        assert [int(k.response) for k in kps] == sc.tolist()
        total += len(kps)
    assert total > 0


def test_fast9_on_reference_sized_cells():
    # the extractor calls cv::FAST on ~(35+6)^2 cells: many tiny images whose borders matter (a 3-px rim is skipped)
    img = synth.scene(480, 640, 7)
    det = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    n = 0
    for y0 in range(16, 400, 38):
        for x0 in range(16, 560, 36):
            cell = np.ascontiguousarray(img[y0:y0 + 44, x0:x0 + 42])
            kps = det.detect(cell)
            xs, ys, sc = orbref.fast9(cell, 20)
            assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps] == list(zip(xs.tolist(), ys.tolist(),
                                                                                          sc.tolist()))
            n += len(kps)
    assert n > 100


def test_fast_atan2_is_cv_fastatan2():
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.integers(-200000, 200000, (20000, 2)).astype(np.float32),
                          np.array([[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0], [1, 1], [-1, -1], [5, -5], [-3, 3]],
                                   np.float32)])
    for y, x in pts:
        a, b = orbref.fast_atan2(y, x), cv2.fastAtan2(float(y), float(x))
        assert np.float32(a).tobytes() == np.float32(b).tobytes(), (y, x, a, b)


def test_cv_round_is_lrintf():
    libm = ctypes.CDLL("libm.so.6")
    libm.lrintf.argtypes = [ctypes.c_float]
    libm.lrintf.restype = ctypes.c_long
    L = orbref.lib()
    for v in (0.5, 1.5, 2.5, -0.5, -1.5, 2.4999, 2.5001, -7.5, 1e6 + 0.5, 13.0, -13.49):
        assert L.orbref_cv_round(v) == libm.lrintf(v)


def test_std_sort_perm_sorts_and_is_deterministic():
    rng = np.random.default_rng(1)
    for n in (1, 2, 15, 16, 17, 100, 513):
        k0, k1 = rng.integers(1, 6, n), rng.integers(0, 8, n) * 40  # tie-heavy like (size, UL.x)
        p = orbref.std_sort_perm(k0, k1)
        assert sorted(p.tolist()) == list(range(n))
        keys = list(zip(k0[p].tolist(), k1[p].tolist()))
        assert keys == sorted(keys)
        assert np.array_equal(p, orbref.std_sort_perm(k0, k1))


@pytest.mark.parametrize("nq,nt,proto", [(500, 600, 0), (300, 400, 16), (50, 2, 0)])
def test_knn2_is_bfmatcher_knnmatch(nq, nt, proto):
    # cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2), src/Frame.cc:1293
    q, t = synth.descriptors(nq, 11, proto), synth.descriptors(nt, 12, proto)
    idx1, d1, idx2, d2 = orbref.knn2(q, t)
    mm = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    assert [m[0].trainIdx for m in mm] == idx1.tolist() and [int(m[0].distance) for m in mm] == d1.tolist()
    assert [m[1].trainIdx for m in mm] == idx2.tolist() and [int(m[1].distance) for m in mm] == d2.tolist()


def test_descriptor_distance_is_popcount():
    a, b = synth.descriptors(200, 1), synth.descriptors(200, 2)
    for i in range(200):
        ref = int(np.unpackbits(a[i] ^ b[i]).sum())
        assert orbref.descriptor_distance(a[i], b[i]) == ref
        assert int(cv2.norm(a[i], b[i], cv2.NORM_HAMMING)) == ref


def test_distinctive_descriptor_matches_a_numpy_restatement():
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:407-435): least median of the sorted distance rows,
    element [0.5 * (N - 1)], first row on ties."""
    rng = np.random.default_rng(9)
    assert orbref.distinctive_descriptor(np.zeros((0, 32), np.uint8)) == -1
    for n in (1, 2, 3, 4, 5, 8, 9, 33, 64, 150):
        d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        if n >= 4:
            d[n // 2] = d[0]
        D = np.bitwise_count(d.view(np.uint64)[:, None, :] ^ d.view(np.uint64)[None, :, :]).sum(axis=2)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert orbref.distinctive_descriptor(d) == int(np.argmin(med)), n


def test_cvt_gray_matches_cv2():
    """cv::cvtColor(..., COLOR_*2GRAY) as Tracking::GrabImage* applies it (src/Tracking.cc:1394-1412): the 15-bit fixed
    point restated in the oracle equals cv2 on every channel order, including odd widths."""
    rng = np.random.default_rng(3)
    for (h, w) in ((64, 4096), (37, 101)):
        px3 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        px4 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        assert np.array_equal(orbref.cvt_gray(px3, rgb=False), cv2.cvtColor(px3, cv2.COLOR_BGR2GRAY))
        assert np.array_equal(orbref.cvt_gray(px3, rgb=True), cv2.cvtColor(px3, cv2.COLOR_RGB2GRAY))
        assert np.array_equal(orbref.cvt_gray(px4, rgb=False), cv2.cvtColor(px4, cv2.COLOR_BGRA2GRAY))
        assert np.array_equal(orbref.cvt_gray(px4, rgb=True), cv2.cvtColor(px4, cv2.COLOR_RGBA2GRAY))
    # every (b, g, r) on a coarse lattice + the extremes
    v = np.array(sorted(set(list(range(0, 256, 5)) + [1, 2, 253, 254, 255])), np.uint8)
    grid = np.stack(np.meshgrid(v, v, v, indexing="ij"), axis=-1).reshape(1, -1, 3)
    assert np.array_equal(orbref.cvt_gray(grid), cv2.cvtColor(grid, cv2.COLOR_BGR2GRAY))


def _rectify_maps(h, w, seed, dh=None, dw=None):
    """Smooth EuRoC-like rectification maps + sub-pixel noise, reaching outside the source on every side."""
    rng = np.random.default_rng(seed)
    dh, dw = dh or h, dw or w
    ys, xs = np.mgrid[0:dh, 0:dw].astype(np.float32)
    mapx = (xs * (w / dw) + 3.7 * np.sin(ys / 50.0) + rng.uniform(-0.5, 0.5, (dh, dw)) - 4).astype(np.float32)
    mapy = (ys * (h / dh) + 2.9 * np.cos(xs / 70.0) + rng.uniform(-0.5, 0.5, (dh, dw)) - 3).astype(np.float32)
    mapx[:3, :5] = -40.25   # far outside
    mapy[-2:, -7:] = h + 9.5
    return mapx, mapy


def test_remap_linear_matches_cv2():
    """cv::remap(..., INTER_LINEAR) with CV_32F maps (src/System.cc:293-294): 1/32-pixel coordinates, exact 5-bit
    weights, constant-0 border — the oracle's restatement equals cv2 byte for byte."""
    rng = np.random.default_rng(11)
    for (h, w, dh, dw) in ((480, 752, 480, 752), (120, 161, 97, 203)):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        mapx, mapy = _rectify_maps(h, w, h + w, dh, dw)
        assert np.array_equal(orbref.remap_linear(src, mapx, mapy), cv2.remap(src, mapx, mapy, cv2.INTER_LINEAR))
