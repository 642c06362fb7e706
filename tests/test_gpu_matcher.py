"""Parity of the CUDA Hamming searches (through the orbm C ABI) with the CPU oracle. Integer outputs bit-exact;
float outputs (uRight, depth) compared bit-exactly as well (the documented bar is 1e-4)."""
import numpy as np
import pytest

from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, synth, views
from oracle import orbref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher(gpu):
    return ORBmatcher(0.8, True)


def test_descriptor_distance(matcher):
    a, b = synth.descriptors(5000, 1), synth.descriptors(5000, 2)
    b[:10] = a[:10]
    d = matcher.DescriptorDistanceBatch(a, b)
    ref = np.array([orbref.descriptor_distance(a[i], b[i]) for i in range(len(a))])
    assert np.array_equal(d, ref)


@pytest.mark.parametrize("nq,nt,proto", [(1000, 1000, 0), (3000, 3000, 0), (1000, 1000, 64), (257, 5000, 16),
                                         (5000, 129, 0), (33, 2, 0), (10, 1, 0), (7, 0, 0), (10000, 10000, 0)])
def test_knn2_matches_oracle(matcher, nq, nt, proto):
    q, t = synth.descriptors(nq, 3, proto), synth.descriptors(nt, 4, proto)
    got = matcher.knnMatch2(q, t)
    ref = orbref.knn2(q, t) if nt > 0 else tuple(np.full(nq, -1, np.int32) for _ in range(4))
    for g, r, name in zip(got, ref, ("idx1", "d1", "idx2", "d2")):
        assert np.array_equal(g, r), "%s differs at %s" % (name, np.nonzero(g != r)[0][:10])


@pytest.mark.parametrize("nq,nt,proto", [(6000, 6000, 64), (8192, 4097, 16), (20000, 3000, 0), (1024, 32768, 8)])
def test_knn2_tensor_core_path_equals_popc_path_and_oracle(matcher, monkeypatch, nq, nt, proto):
    """Sizes above the switch-over run as an s8 GEMM on the tensor cores (k_knn2_tc.cu); ORBM_KNN2_TC=0 keeps the call on
    the POPC kernel. Both must give the oracle's words, ties (duplicate rows, exact hits) included."""
    q, t = synth.descriptors(nq, 11, proto), synth.descriptors(nt, 12, proto)
    q[: min(nq, nt) // 2] = t[: min(nq, nt) // 2]  # exact hits (d = 0), duplicated among the prototypes
    tc = matcher.knnMatch2(q, t)
    monkeypatch.setenv("ORBM_KNN2_TC", "0")
    popc = matcher.knnMatch2(q, t)
    monkeypatch.delenv("ORBM_KNN2_TC")
    ref = orbref.knn2(q, t)
    for a, b, r, name in zip(tc, popc, ref, ("idx1", "d1", "idx2", "d2")):
        assert np.array_equal(a, r), "tensor-core %s differs at %s" % (name, np.nonzero(a != r)[0][:10])
        assert np.array_equal(b, r), "POPC %s differs at %s" % (name, np.nonzero(b != r)[0][:10])


def test_knn2_tensor_core_fallback_form_equals_oracle():
    """k_knn2_tc_ts (256 queries per CTA, query tiles in tensor memory) is the tensor-core form that normally runs; the
    first form (k_knn2_tc: query tile in shared memory, 128 x 256 tiles) is kept behind ORBM_KNN2_TS=0. The switch is
    read once per process, hence the subprocess. Both must give the oracle's words."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np\n"
        "from orb_slam3_fast_b200 import ORBmatcher, synth\n"
        "from oracle import orbref\n"
        "m = ORBmatcher()\n"
        "for nq, nt, proto in ((6000, 6000, 64), (8192, 4097, 16), (2000, 40000, 0)):\n"
        "    q, t = synth.descriptors(nq, 21, proto), synth.descriptors(nt, 22, proto)\n"
        "    q[: min(nq, nt) // 2] = t[: min(nq, nt) // 2]\n"
        "    got, ref = m.knnMatch2(q, t), orbref.knn2(q, t)\n"
        "    assert all(np.array_equal(g, r) for g, r in zip(got, ref)), (nq, nt, proto)\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for ts in ("0", "1"):
        env = dict(os.environ, ORBM_KNN2_TS=ts, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode == 0 and "ok" in r.stdout, "ORBM_KNN2_TS=%s\n%s\n%s" % (ts, r.stdout[-2000:], r.stderr[-2000:])


def test_knn2_agrees_with_cv2_bfmatcher(matcher):
    cv2 = pytest.importorskip("cv2")
    q, t = synth.descriptors(800, 5, 32), synth.descriptors(900, 6, 32)  # tie-heavy
    idx1, d1, idx2, d2 = matcher.knnMatch2(q, t)
    mm = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    assert [m[0].trainIdx for m in mm] == idx1.tolist() and [m[1].trainIdx for m in mm] == idx2.tolist()
    assert [int(m[0].distance) for m in mm] == d1.tolist() and [int(m[1].distance) for m in mm] == d2.tolist()


def _stereo_case(w, h, nf, seed, kind="scene"):
    left, right, _ = synth.stereo_pair(h, w, seed, kind=kind)
    exl, exr = ORBextractor(nf), ORBextractor(nf)
    rl, rr = orbref.Extractor(nf), orbref.Extractor(nf)
    _, kl, dl = exl(left)
    _, kr, dr = exr(right)
    _, kl_r, dl_r = rl(left, (0, 0))
    _, kr_r, dr_r = rr(right, (0, 0))
    assert np.array_equal(kl, kl_r) and np.array_equal(kr, kr_r) and np.array_equal(dl, dl_r)
    return (exl, exr, rl, rr), (kl, dl, kr, dr)


@pytest.mark.parametrize("warps", ["4", "8"])
def test_stereo_match_both_cta_shapes_equal_oracle(matcher, monkeypatch, warps):
    """k_stereo_match runs with 8 warps per CTA in small launches (latency) and with 4 in launches that fill the chip
    (throughput); ORBX_STEREO_WARPS forces one form. Both must give the oracle's words on the same pair."""
    monkeypatch.setenv("ORBX_STEREO_WARPS", warps)
    (exl, exr, rl, rr), (kl, dl, kr, dr) = _stereo_case(640, 480, 1200, 7, kind="noise_blur")
    mbf, mb = float(np.float32(435.2 * 0.11)), float(np.float32(0.11))
    nm, ur, dp = matcher.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, mbf, mb)
    nm_r, ur_r, dp_r = orbref.stereo_match(rl, rr, kl, dl, kr, dr, mbf, mb)
    assert nm == nm_r and nm > 50 and ur.tobytes() == ur_r.tobytes() and dp.tobytes() == dp_r.tobytes()


@pytest.mark.parametrize("w,h,nf,seed", [(752, 480, 1200, 0), (640, 480, 1000, 1), (1280, 720, 2000, 2)])
def test_stereo_matches_oracle(matcher, w, h, nf, seed):
    (exl, exr, rl, rr), (kl, dl, kr, dr) = _stereo_case(w, h, nf, seed)
    fx = 435.2
    mbf, mb = np.float32(fx * 0.11), np.float32(0.11)
    n, ur, dp = matcher.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, float(mbf), float(mb))
    n_r, ur_r, dp_r = orbref.stereo_match(rl, rr, kl, dl, kr, dr, float(mbf), float(mb))
    assert n_r > 50, "degenerate test: oracle found only %d matches" % n_r
    assert n == n_r
    assert np.array_equal(ur, ur_r), "uRight differs at %s" % np.nonzero(ur != ur_r)[0][:10]
    assert np.array_equal(dp, dp_r)


def test_stereo_no_matches_is_defined(matcher):
    exl, exr = ORBextractor(1000), ORBextractor(1000)
    _, kl, dl = exl(synth.scene(480, 640, 1))
    _, kr, dr = exr(synth.scene(480, 640, 99))  # unrelated image
    n, ur, dp = matcher.ComputeStereoMatches(exl, exr, kl, dl, kr[:0], dr[:0], 40.0, 0.1)
    assert n == 0 and (ur == -1).all() and (dp == -1).all()


def _frame_views(kps, desc, w, h, scale, u_right, occupied):
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    off, items = views.assign_features_to_grid(kps, 0.0, 0.0, inv_w, inv_h)
    fv = views.make_frame_view(kps, desc, u_right, occupied, off, items, 0.0, 0.0, inv_w, inv_h, scale)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    fr = orbref.make_frame_view(kps, desc, u_right, occupied, g, keep, scale)
    return fv, fr


@pytest.mark.parametrize("m,th,stereo,seed", [(10000, 1.0, True, 0), (3000, 3.0, False, 1), (2000, 15.0, True, 2)])
def test_search_by_projection_map(matcher, m, th, stereo, seed):
    w, h = 640, 480
    ex = ORBextractor(1200)
    _, kps, desc = ex(synth.scene(h, w, seed))
    rng = np.random.default_rng(seed)
    ur = np.where(rng.random(len(kps)) < 0.7, kps["x"] - rng.uniform(1, 40, len(kps)), -1).astype(np.float32)
    occ = (rng.random(len(kps)) < 0.1).astype(np.uint8)
    fv, fr = _frame_views(kps, desc, w, h, ex.GetScaleFactors(), ur if stereo else None, occ)
    mp = synth.local_map(kps, desc, m, w, h, 8, seed)
    n, assign = matcher.SearchByProjection(fv, views.make_mappoints(**mp), th, True, 15.0)
    n_r, assign_r = orbref.search_by_projection_map(fr, orbref.make_mappoints(**mp), th, 0.8, True, 15.0)
    assert n_r > 100
    assert n == n_r
    assert np.array_equal(assign, assign_r), "assign differs at %s" % np.nonzero(assign != assign_r)[0][:10]


@pytest.mark.parametrize("kind,n_extra,seed", [("scene", 0, 0), ("uniform_noise", 300, 1), ("scene", 50, 2)])
def test_assign_features_to_grid(matcher, kind, n_extra, seed):
    # Frame::AssignFeaturesToGrid (src/Frame.cc:520-547): device CSR == oracle == the host mirror, including keypoints
    # that fall outside the grid (PosInGrid false) and cells holding many keypoints
    w, h = 640, 480
    _, kps, _ = ORBextractor(1200)(synth.make(kind, h, w, seed))
    rng = np.random.default_rng(seed)
    if n_extra:
        extra = np.zeros(n_extra, kps.dtype)
        extra["x"] = rng.uniform(-30, w + 30, n_extra).astype(np.float32)   # some land outside [0, 64) x [0, 48)
        extra["y"] = rng.uniform(-30, h + 30, n_extra).astype(np.float32)
        extra[: n_extra // 3] = kps[: n_extra // 3]                          # duplicates: several indices per cell
        kps = np.concatenate([kps, extra])[rng.permutation(len(kps) + n_extra)]
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    for (mx, my) in ((0.0, 0.0), (-7.5, 3.25)):
        off, items = matcher.AssignFeaturesToGrid(kps, mx, my, inv_w, inv_h)
        off_r, items_r = orbref.build_grid(kps, mx, my, inv_w, inv_h)
        off_h, items_h = views.assign_features_to_grid(kps, mx, my, inv_w, inv_h)
        assert np.array_equal(off, off_r) and np.array_equal(items, items_r[: off_r[-1]])
        assert np.array_equal(off, off_h) and np.array_equal(items, items_h[: off_h[-1]])
    off, items = matcher.AssignFeaturesToGrid(kps[:0], 0.0, 0.0, inv_w, inv_h)
    assert (off == 0).all() and len(items) == 0


@pytest.mark.parametrize("m,th,stereo,seed", [(10000, 1.0, True, 0), (3000, 3.0, False, 1)])
def test_search_by_projection_resident_frame(matcher, m, th, stereo, seed):
    # configs[3]: extract, then SearchByProjection against a 10k-point local map without the frame leaving the device
    w, h = 640, 480
    ex = ORBextractor(1200, max_batch=4)
    imgs = np.stack([synth.scene(h, w, seed + k) for k in range(3)])
    n_all, _, kps_all, desc_all = ex.extract_batch(imgs, (0, 0))
    f = 1
    kps, desc = kps_all[f, : n_all[f]], desc_all[f, : n_all[f]]
    rng = np.random.default_rng(seed)
    ur = np.where(rng.random(len(kps)) < 0.7, kps["x"] - rng.uniform(1, 40, len(kps)), -1).astype(np.float32)
    occ = (rng.random(len(kps)) < 0.1).astype(np.uint8)
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    _, fr = _frame_views(kps, desc, w, h, ex.GetScaleFactors(), ur if stereo else None, occ)
    mp = synth.local_map(kps, desc, m, w, h, 8, seed)
    n, assign = matcher.SearchByProjectionResident(ex, f, len(kps), views.make_mappoints(**mp), (0.0, 0.0, inv_w, inv_h),
                                                   ur if stereo else None, occ, th, True, 15.0)
    n_r, assign_r = orbref.search_by_projection_map(fr, orbref.make_mappoints(**mp), th, 0.8, True, 15.0)
    assert n_r > 100 and n == n_r
    assert np.array_equal(assign, assign_r), "assign differs at %s" % np.nonzero(assign != assign_r)[0][:10]
    with pytest.raises(Exception):
        matcher.SearchByProjectionResident(ex, 7, len(kps), views.make_mappoints(**mp), (0.0, 0.0, inv_w, inv_h))


@pytest.mark.parametrize("m,th,stereo,check,seed", [(1200, 7.0, True, True, 0), (1200, 15.0, False, True, 1),
                                                    (4000, 7.0, True, False, 2)])
def test_search_by_projection_frame(gpu, m, th, stereo, check, seed):
    w, h = 752, 480
    ex = ORBextractor(1200)
    _, kps, desc = ex(synth.scene(h, w, seed))
    rng = np.random.default_rng(seed)
    ur = np.where(rng.random(len(kps)) < 0.7, kps["x"] - rng.uniform(1, 40, len(kps)), -1).astype(np.float32)
    occ = np.zeros(len(kps), np.uint8)
    fv, fr = _frame_views(kps, desc, w, h, ex.GetScaleFactors(), ur if stereo else None, occ)
    pts = synth.projected_points(kps, desc, m, w, h, 8, ex.GetScaleFactors(), seed, th=th, stereo=stereo)
    mt = ORBmatcher(0.9, check)
    n, assign = mt.SearchByProjectionProjected(fv, views.make_projected(**pts), 100)
    n_r, assign_r = orbref.search_by_projection_frame(fr, orbref.make_projected(**pts), 100, check)
    assert n_r > 100
    assert n == n_r
    assert np.array_equal(assign, assign_r)
    # the same search on the frame as the extractor left it on the device (grid built there too)
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    n2, assign2 = mt.SearchByProjectionProjectedResident(ex, 0, len(kps), views.make_projected(**pts),
                                                         (0.0, 0.0, inv_w, inv_h), ur if stereo else None, occ, 100)
    assert n2 == n_r and np.array_equal(assign2, assign_r)


@pytest.mark.parametrize("m,th,seed", [(1200, 7.0, 3), (4000, 15.0, 4), (300, 1.0, 5)])
def test_search_by_projection_frame_decisions(gpu, m, th, seed):
    """orbm_search_by_projection_frame_decisions — one camera of SearchByProjection(CurrentFrame, LastFrame) on a
    two-camera frame: the keypoint every point takes under the serial order dependence (occupied keypoints, points
    without observations that do not block) and |GetFeaturesInArea| per window, against the oracle's serial loop."""
    w, h = 752, 480
    ex = ORBextractor(1200)
    _, kps, desc = ex(synth.scene(h, w, seed))
    rng = np.random.default_rng(seed)
    occ = (rng.random(len(kps)) < 0.15).astype(np.uint8)
    fv, fr = _frame_views(kps, desc, w, h, ex.GetScaleFactors(), None, occ)
    pts = synth.projected_points(kps, desc, m, w, h, 8, ex.GetScaleFactors(), seed, th=th, stereo=False)
    pts["has_obs"] = (rng.random(m) < 0.7).astype(np.uint8)
    # a fifth of the windows far from every keypoint cluster
    far = rng.random(m) < 0.2
    pts["u"] = np.where(far, rng.uniform(0, w, m), pts["u"]).astype(np.float32)
    pts["v"] = np.where(far, rng.uniform(0, h, m), pts["v"]).astype(np.float32)
    mt = ORBmatcher(0.9, True)
    dec, win = mt.SearchByProjectionProjectedDecisions(fv, views.make_projected(**pts), 100)
    dec_r, win_r = orbref.search_by_projection_frame_decisions(fr, orbref.make_projected(**pts), 100)
    assert (dec_r >= 0).sum() > m // 8 and (win_r == 0).any()
    assert np.array_equal(win, win_r), np.nonzero(win != win_r)[0][:10]
    assert np.array_equal(dec, dec_r), np.nonzero(dec != dec_r)[0][:10]


@pytest.mark.parametrize("only_stereo,coarse,check,seed", [(False, False, True, 0), (True, False, True, 1),
                                                           (False, True, False, 2)])
def test_search_for_triangulation(gpu, only_stereo, coarse, check, seed):
    w, h = 752, 480
    left, right, _ = synth.stereo_pair(h, w, seed, d_min=5, d_max=30)
    e1, e2 = ORBextractor(1500), ORBextractor(1500)
    _, k1, d1 = e1(left)
    _, k2, d2 = e2(right)
    rng = np.random.default_rng(seed)
    # vocabulary nodes: make corresponding features (similar descriptors) likely to share a node by hashing 10 bits
    def featvec(desc):
        node_of = (desc[:, 0].astype(np.int64) >> 3) * 32 + (desc[:, 1].astype(np.int64) >> 3)
        ids, inv = np.unique(node_of, return_inverse=True)
        order = np.argsort(inv, kind="stable")
        offsets = np.zeros(len(ids) + 1, np.int32)
        offsets[1:] = np.cumsum(np.bincount(inv, minlength=len(ids)))
        return ids.astype(np.uint32), offsets, order.astype(np.uint32)
    sf, s2 = e1.GetScaleFactors(), e1.GetScaleSigmaSquares()
    args = []
    for k, d in ((k1, d1), (k2, d2)):
        ids, off, idx = featvec(d)
        ur = np.where(rng.random(len(k)) < 0.5, k["x"] - 10, -1).astype(np.float32)
        hm = (rng.random(len(k)) < 0.2).astype(np.uint8)
        args.append((k, d, ur, hm, ids, off, idx, sf, s2))
    # pure horizontal translation: F = [t]x up to scale => epipolar lines are image rows
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
    ep = (1e6, 240.0)
    mt = ORBmatcher(0.6, check)
    n, m12, pairs = mt.SearchForTriangulation(views.make_keyframe_view(*args[0]), views.make_keyframe_view(*args[1]),
                                              F12, ep, only_stereo, coarse)
    n_r, m12_r = orbref.search_for_triangulation(orbref.make_keyframe_view(*args[0]),
                                                 orbref.make_keyframe_view(*args[1]), F12, ep, only_stereo, coarse,
                                                 check)
    assert n_r > 20, "degenerate test: %d" % n_r
    assert n == n_r and np.array_equal(m12, m12_r)
    assert len(pairs) == n


def test_stereo_frames_batch_matches_per_pair_oracle(matcher):
    """Fused host-facing batch (extract x2 + ComputeStereoMatches, pipelined over lanes) == oracle pair by pair."""
    w, h, nf = 752, 480, 1200
    pairs = [synth.stereo_pair(h, w, s) for s in range(5)]
    L = np.stack([p[0] for p in pairs])
    R = np.stack([p[1] for p in pairs])
    exl, exr = ORBextractor(nf, max_batch=2), ORBextractor(nf, max_batch=2)  # 5 pairs -> groups 2 + 2 + 1
    mbf, mb = float(np.float32(435.2 * 0.11)), float(np.float32(0.11))
    out = matcher.StereoFramesBatch(exl, exr, L, R, mbf, mb)
    rl, rr = orbref.Extractor(nf), orbref.Extractor(nf)
    for i in range(len(pairs)):
        _, kl, dl = rl(L[i], (0, 0))
        _, kr, dr = rr(R[i], (0, 0))
        n_r, ur_r, dp_r = orbref.stereo_match(rl, rr, kl, dl, kr, dr, mbf, mb)
        nl, nr = out["n_l"][i], out["n_r"][i]
        assert nl == len(kl) and nr == len(kr)
        assert np.array_equal(out["kps_l"][i, :nl], kl) and np.array_equal(out["desc_l"][i, :nl], dl)
        assert np.array_equal(out["kps_r"][i, :nr], kr) and np.array_equal(out["desc_r"][i, :nr], dr)
        assert out["n_matched"][i] == n_r
        assert np.array_equal(out["u_right"][i, :nl], ur_r) and np.array_equal(out["depth"][i, :nl], dp_r)


def test_compute_distinctive_descriptors(matcher):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:372-441) for a batch of map points: list sizes 0, 1, 2,
    odd / even, > 32 and > 256 observations, duplicated descriptors (median ties -> the first row wins)."""
    rng = np.random.default_rng(5)
    sizes = [0, 1, 2, 3, 4, 7, 8, 31, 32, 33, 64, 100, 300] + list(rng.integers(2, 40, 200))
    lists = []
    for n in sizes:
        base = synth.descriptors(1, int(rng.integers(1 << 30)))[0]
        d = np.stack([synth.flip_bits(base[None], [int(rng.integers(0, 60))], rng)[0] for _ in range(n)]) if n else \
            np.zeros((0, 32), np.uint8)
        if n >= 4:
            d[n // 2] = d[0]            # exact duplicates
            d[n - 1] = d[1]
        lists.append(d)
    best = matcher.ComputeDistinctiveDescriptors(lists)
    ref = np.array([orbref.distinctive_descriptor(d) for d in lists], np.int32)
    assert np.array_equal(best, ref), np.nonzero(best != ref)[0][:10]
    assert len(matcher.ComputeDistinctiveDescriptors([])) == 0


@pytest.mark.parametrize("nnratio,check,seed", [(0.7, True, 0), (0.9, False, 1), (0.75, True, 2)])
def test_search_by_bow_keyframe_frame(gpu, nnratio, check, seed):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:230-404): per shared vocabulary node the
    KeyFrame features that carry a MapPoint take the best still-free frame feature (ratio test, TH_LOW), then the
    rotation histogram. Frame = the right image of a stereo pair, so corresponding descriptors are close."""
    w, h = 752, 480
    left, right, _ = synth.stereo_pair(h, w, seed, d_min=3, d_max=20)
    e1, e2 = ORBextractor(1500), ORBextractor(1500)
    _, k1, d1 = e1(left)
    _, k2, d2 = e2(right)
    rng = np.random.default_rng(seed)

    def featvec(desc, bits):
        # coarse hash of the descriptor = vocabulary node: few nodes (many features per node) stresses the greedy part
        node_of = (desc[:, 0].astype(np.int64) >> (8 - bits)) * 7 + 3
        ids, inv = np.unique(node_of, return_inverse=True)
        order = np.argsort(inv, kind="stable")
        offsets = np.zeros(len(ids) + 1, np.int32)
        offsets[1:] = np.cumsum(np.bincount(inv, minlength=len(ids)))
        return ids.astype(np.uint32), offsets, order.astype(np.uint32)
    sf, s2 = e1.GetScaleFactors(), e1.GetScaleSigmaSquares()
    for bits in (3, 6):
        views_g, views_r = [], []
        for k, d in ((k1, d1), (k2, d2)):
            ids, off, idx = featvec(d, bits)
            ur = np.full(len(k), -1, np.float32)
            hm = (rng.random(len(k)) < 0.6).astype(np.uint8)   # KeyFrame features that carry a good MapPoint
            a = (k, d, ur, hm, ids, off, idx, sf, s2)
            views_g.append(views.make_keyframe_view(*a))
            views_r.append(orbref.make_keyframe_view(*a))
        mt = ORBmatcher(nnratio, check)
        n, mf = mt.SearchByBoW(views_g[0], views_g[1])
        n_r, mf_r = orbref.search_by_bow(views_r[0], views_r[1], nnratio, check)
        assert n_r > 20, "degenerate test: %d matches" % n_r
        assert n == n_r and np.array_equal(mf, mf_r), np.nonzero(mf != mf_r)[0][:10]
        # the KeyFrame-KeyFrame form (:766-884): MapPoints required on both sides, strict < TH_LOW, matches by kf1 index
        n2, m12 = mt.SearchByBoWKeyFrames(views_g[0], views_g[1])
        n2_r, m12_r = orbref.search_by_bow_kf(views_r[0], views_r[1], nnratio, check)
        assert n2_r > 10, "degenerate test: %d matches" % n2_r
        assert n2 == n2_r and np.array_equal(m12, m12_r), np.nonzero(m12 != m12_r)[0][:10]


@pytest.mark.parametrize("seed", [5, 6])
def test_triangulation_candidates(gpu, seed):
    """orbm_triangulation_candidates: the descriptor part of SearchForTriangulation (src/ORBmatcher.cc:973-988) for rigs
    whose epipolar test stays with the caller's camera objects — CSR of (idx2, distance <= TH_LOW) per kf1 feature in
    scan order == oracle; a too small buffer is reported with the required size."""
    from test_oracle_matchers_vs_reference_source import _triangulation_case
    v1r, v2r = _triangulation_case(seed)   # the view structs are the ABI's: the oracle's holders serve both sides
    off_r, idx_r, dist_r = orbref.triangulation_candidates(v1r, v2r)
    mt = ORBmatcher(0.6, True)
    off, idx, dist = mt.TriangulationCandidates(v1r, v2r)
    assert len(idx_r) > 100
    assert np.array_equal(off, off_r) and np.array_equal(idx, idx_r) and np.array_equal(dist, dist_r)
    off2, idx2, dist2 = mt.TriangulationCandidates(v1r, v2r, cap=7)     # retried with the reported total
    assert np.array_equal(off2, off_r) and np.array_equal(idx2, idx_r) and np.array_equal(dist2, dist_r)


@pytest.mark.parametrize("nnratio,check,seed", [(0.7, True, 1), (0.9, False, 2), (0.6, True, 3)])
def test_search_by_bow_two_camera_frame(gpu, nnratio, check, seed):
    """SearchByBoW(KeyFrame*, Frame&, ...) on a two-camera Frame (F.Nleft != -1, src/ORBmatcher.cc:274-365): left /
    right bests kept apart, the right best accepted without a ratio test inside the left one's TH_LOW gate. The oracle
    is pinned to the reference's own method (tests/test_oracle_matchers_vs_reference_source.py)."""
    from test_oracle_matchers import _keyframes, _two_cameras
    k1, k2 = _keyframes(seed, n=600, n_nodes=9)
    k2, nl = _two_cameras(k2, seed, n_nodes=9)
    sf = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    args = [(k["kps"], k["desc"], None, k["hm"], k["ids"], k["off"], k["idx"], sf, sf * sf) for k in (k1, k2)]
    vg = [views.make_keyframe_view(*a) for a in args]
    vr = [orbref.make_keyframe_view(*a) for a in args]
    mt = ORBmatcher(nnratio, check)
    for n_left in (nl, len(k2["kps"]), 0, nl // 2):
        n, mf = mt.SearchByBoWTwoCameras(vg[0], vg[1], n_left)
        n_r, mf_r = orbref.search_by_bow_fisheye(vr[0], vr[1], n_left, nnratio, check)
        assert n == n_r and np.array_equal(mf, mf_r), (n_left, n, n_r, np.nonzero(mf != mf_r)[0][:10])
        if n_left == nl:
            assert (mf_r[:nl] >= 0).sum() > 20 and (mf_r[nl:] >= 0).sum() > 20, "degenerate test"


@pytest.mark.parametrize("m,th,stereo,seed", [(3000, 3.0, True, 0), (1500, 2.5, False, 1), (6000, 4.0, True, 2)])
def test_fuse_match(matcher, m, th, stereo, seed):
    """The matching loop of ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>, th) (src/ORBmatcher.cc:1194-1257): window
    from KeyFrame::GetFeaturesInArea, level window [L-1, L], chi-square gate (7.8 stereo / 5.99 mono), best distance
    with the first keypoint winning ties."""
    w, h = 752, 480
    ex = ORBextractor(1200)
    _, kps, desc = ex(synth.scene(h, w, seed + 20))
    rng = np.random.default_rng(seed)
    n = len(kps)
    ur_kf = np.where(rng.random(n) < 0.6, kps["x"] - rng.uniform(1, 40, n), -1).astype(np.float32) if stereo else None
    occ = np.zeros(n, np.uint8)
    fv, fr = _frame_views(kps, desc, w, h, ex.GetScaleFactors(), ur_kf, occ)
    inv_sigma2 = 1.0 / np.asarray(ex.GetScaleSigmaSquares(), np.float32)
    # projected MapPoints: most sit within a pixel or two of a keypoint (so that the chi-square gate decides), with the
    # keypoint's level or a neighbouring one and a descriptor a few bits away; duplicates force distance ties
    src = rng.integers(0, n, m)
    near = rng.random(m) < 0.85
    u = np.where(near, kps["x"][src] + rng.normal(0, 1.2, m), rng.uniform(0, w, m)).astype(np.float32)
    v = np.where(near, kps["y"][src] + rng.normal(0, 1.2, m), rng.uniform(0, h, m)).astype(np.float32)
    level = np.clip(kps["octave"][src] + rng.integers(-1, 2, m), 0, 7).astype(np.int32)
    radius = (np.float32(th) * np.asarray(ex.GetScaleFactors(), np.float32)[level]).astype(np.float32)
    d = synth.flip_bits(desc[src], rng.integers(0, 60, m), rng)
    d[::7] = desc[src[::7]]
    base_ur = ur_kf[src] if stereo else np.zeros(m, np.float32)
    u_r = np.where(base_ur >= 0, base_ur + rng.normal(0, 1.0, m), u - 5).astype(np.float32)
    pts = dict(u=u, v=v, u_right=u_r if stereo else None, radius=radius, min_level=level - 1, max_level=level,
               angle=np.zeros(m, np.float32), has_obs=np.zeros(m, np.uint8), desc=d)
    bi, bd = matcher.FuseMatch(fv, inv_sigma2, views.make_projected(**pts))
    bi_r, bd_r = orbref.fuse_match(fr, inv_sigma2, orbref.make_projected(**pts))
    assert (bi_r >= 0).sum() > m // 4 and ((bi_r < 0) & near).sum() > 0, "degenerate test"
    assert np.array_equal(bi, bi_r), np.nonzero(bi != bi_r)[0][:10]
    assert np.array_equal(bd, bd_r)
    # Fuse(KeyFrame*, Sim3f&, ...) (:1277-1390): the same loop without the reprojection gate
    bi2, bd2 = matcher.FuseMatch(fv, inv_sigma2, views.make_projected(**pts), chi2_gate=False)
    bi2_r, bd2_r = orbref.fuse_match(fr, inv_sigma2, orbref.make_projected(**pts), chi2_gate=False)
    assert np.array_equal(bi2, bi2_r) and np.array_equal(bd2, bd2_r) and (bi2_r >= 0).sum() > (bi_r >= 0).sum()


@pytest.mark.parametrize("window,nnratio,check,seed", [(100, 0.9, True, 0), (30, 0.9, False, 1), (100, 0.8, True, 2)])
def test_search_for_initialization(gpu, window, nnratio, check, seed):
    """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:618-764) in the serial order of its loop: level-0
    keypoints of F1 against the level-0 keypoints of F2 inside a window around vbPrevMatched, the running
    vMatchedDistance / eviction logic, rotation histogram."""
    w, h = 640, 480
    img1 = synth.scene(h, w, seed + 70)
    img2 = np.roll(img1, (3, -5), axis=(0, 1))                 # a small camera motion
    e1, e2 = ORBextractor(2000), ORBextractor(2000)           # the initialiser uses 5 x nFeatures
    _, k1, d1 = e1(img1)
    _, k2, d2 = e2(img2)
    occ1, occ2 = np.zeros(len(k1), np.uint8), np.zeros(len(k2), np.uint8)
    f1, r1 = _frame_views(k1, d1, w, h, e1.GetScaleFactors(), None, occ1)
    f2, r2 = _frame_views(k2, d2, w, h, e2.GetScaleFactors(), None, occ2)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)   # vbPrevMatched starts as F1's keypoints (:2414)
    mt = ORBmatcher(nnratio, check)
    n, m12 = mt.SearchForInitialization(f1, f2, prev, window)
    n_r, m12_r = orbref.search_for_initialization(r1, r2, prev, window, nnratio, check)
    assert n_r > 50, "degenerate test: %d matches" % n_r
    assert n == n_r and np.array_equal(m12, m12_r), np.nonzero(m12 != m12_r)[0][:10]
    assert (m12[k1["octave"] > 0] == -1).all()


@pytest.mark.parametrize("k,depth,levelsup,seed", [(10, 4, 2, 0), (10, 3, 4, 1), (6, 5, 4, 2)])
def test_bow_transform(matcher, k, depth, levelsup, seed):
    """The per-feature part of Frame::ComputeBoW (src/Frame.cc:846-851; DBoW2 TemplatedVocabulary::transform,
    TemplatedVocabulary.h:1218-1262) on a synthetic vocabulary: word, weight and FeatureVector node of every feature,
    including distance ties between siblings, ragged nodes and levelsup >= depth (node = root)."""
    voc = synth.vocabulary(k, depth, seed)
    rng = np.random.default_rng(seed)
    n_nodes = len(voc["descriptors"])
    src = rng.integers(1, n_nodes, 3000)
    feats = synth.flip_bits(voc["descriptors"][src], rng.integers(0, 40, len(src)), rng)
    feats[::11] = voc["descriptors"][src[::11]]           # exact copies of node descriptors
    feats = np.concatenate([feats, synth.descriptors(500, seed + 5)])
    matcher.SetVocabulary(views.make_vocabulary(**voc))
    w, wt, nd = matcher.BowTransform(feats, levelsup)
    w_r, wt_r, nd_r = orbref.bow_transform(orbref.make_vocabulary(**voc), feats, levelsup)
    assert np.array_equal(w, w_r) and np.array_equal(nd, nd_r) and np.array_equal(wt, wt_r)
    assert len(np.unique(w_r)) > 50
    bow, fv = views.assemble_bow(w, wt, nd)
    assert abs(sum(bow.values()) - 1.0) < 1e-9 and sum(len(x) for x in fv.values()) == int((wt > 0).sum())
    if depth - levelsup <= 0:
        assert (nd == 0).all()


@pytest.mark.parametrize("seed,th,far,sizes", [(0, 1.0, True, (700, 650, 3000)), (1, 3.0, False, (700, 650, 3000)),
                                               (2, 6.0, True, (1200, 1100, 10000)), (3, 1.0, False, (40, 0, 200))])
def test_search_by_projection_map_fisheye(matcher, seed, th, far, sizes):
    """The two-camera form (Nleft != -1) of SearchByProjection(Frame&, vector<MapPoint*>), src/ORBmatcher.cc:42-221 with the
    right-camera twin :148-217: left and right searches per point in serial order, stereo partners written
    unconditionally, slots re-opening when a point without observations replaces one with. Every slot of mvpMapPoints
    and the match count must equal the oracle's (which equals the reference's own method, CPU tests)."""
    nl, nr, m = sizes
    fr, mp, mpr = synth.fisheye_case(nl, nr, m, seed=seed) if nr else _left_only_case(nl, m, seed)
    sf = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    inv_w, inv_h = np.float32(64) / np.float32(640), np.float32(48) / np.float32(480)
    args = (fr["kps_left"], fr["kps_right"], fr["desc"], fr["occupied"], 0.0, 0.0, inv_w, inv_h, fr["left_to_right"],
            fr["right_to_left"], sf)
    n_o, a_o = orbref.search_by_projection_map_fisheye(orbref.make_fisheye_view(*args), orbref.make_mappoints(**mp),
                                                       orbref.make_mappoints_right(**mpr), th, 0.8, far, 15.0)
    n_g, a_g = matcher.SearchByProjectionFisheye(views.make_fisheye_view(*args), views.make_mappoints(**mp),
                                                 views.make_mappoints_right(**mpr), th, far, 15.0)
    assert m < 1000 or n_o > 300, "degenerate test"
    assert n_g == n_o and np.array_equal(a_g, a_o), np.nonzero(a_g != a_o)[0][:10]


def _left_only_case(nl, m, seed):
    fr, mp, mpr = synth.fisheye_case(nl, max(nl, 1), m, seed=seed)
    fr["kps_right"], fr["right_to_left"] = fr["kps_right"][:0], fr["right_to_left"][:0]
    fr["desc"], fr["occupied"] = fr["desc"][:nl], fr["occupied"][:nl]
    fr["left_to_right"] = np.full(nl, -1, np.int32)
    return fr, mp, mpr
