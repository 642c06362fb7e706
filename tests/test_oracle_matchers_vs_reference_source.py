"""The oracle's matcher restatements against the reference's OWN src/ORBmatcher.cc. oracle/_ref/liborbref_matcher_src.so
is that file (all 1975 lines, plus Thirdparty/DBoW2/DBoW2/FeatureVector.cpp) compiled where it lies (oracle/Makefile,
target `ref`) against a stand-in world (oracle/ref_stubs/matcher_world.h): Frame / KeyFrame / MapPoint as plain data with
the members the matcher touches (their real headers need Eigen, Sophus, g2o and boost); their grid functions
(AssignFeaturesToGrid, PosInGrid, both GetFeaturesInArea) and Frame::ComputeStereoMatches are the reference's own text
too, cut out of src/Frame.cc / src/KeyFrame.cc by signature and piped to the compiler; TBB run serially, an orthographic stand-in camera whose epipolarConstrain evaluates
Pinhole.cpp:136-148 on a supplied F12. The matching loops, thresholds, ratio tests, rotation histograms and bookkeeping
that run are the reference's code. CPU only; skipped where the reference tree was not available at build time."""
import numpy as np
import pytest

from orb_slam3_fast_b200 import synth
from oracle import orbref, refsrc
from test_oracle_matchers import _keyframes, _two_cameras, _view

pytestmark = pytest.mark.skipif(not refsrc.matcher_available(),
                                reason="oracle/_ref not built (no /root/reference at build time)")
f32 = np.float32


def test_descriptor_distance():
    """ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1959-1973 (the SWAR popcount)."""
    a, b = synth.descriptors(500, 1), synth.descriptors(500, 2)
    b[:20] = a[:20]
    b[20:40] = ~a[20:40]
    for i in range(500):
        assert refsrc.descriptor_distance(a[i], b[i]) == orbref.descriptor_distance(a[i], b[i])


def _frame(n, w, h, seed, stereo, occupied_rate=0.1):
    rng = np.random.default_rng(seed)
    kps = np.zeros(n, synth.KP_DTYPE)
    kps["x"], kps["y"] = rng.uniform(0, w, n).astype(f32), rng.uniform(0, h, n).astype(f32)
    kps["octave"] = rng.integers(0, 8, n)
    kps["angle"] = rng.uniform(0, 360, n).astype(f32)
    desc = synth.descriptors(n, seed)
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    ur = np.where(rng.random(n) < 0.7, kps["x"] - rng.uniform(1, 40, n), -1).astype(f32) if stereo else None
    occ = (rng.random(n) < occupied_rate).astype(np.uint8)
    return kps, desc, sf, orbref.make_frame_view(kps, desc, ur, occ, g, keep, sf)


@pytest.mark.parametrize("stereo,th,nnratio,far,seed", [(True, 1.0, 0.8, True, 3), (False, 3.0, 0.8, False, 4),
                                                        (True, 5.0, 0.6, True, 5), (True, 15.0, 0.9, False, 6)])
def test_search_by_projection_map(stereo, th, nnratio, far, seed):
    """SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints), :42-221 — row a13."""
    kps, desc, sf, fv = _frame(600, 640, 480, seed, stereo)
    mp = synth.local_map(kps, desc, 2000, 640, 480, 8, seed)
    mps = orbref.make_mappoints(**mp)
    n_o, a_o = orbref.search_by_projection_map(fv, mps, th, nnratio, far, 15.0)
    n_r, a_r = refsrc.search_by_projection_map(fv, mps, th, nnratio, far, 15.0)
    assert n_o > 30
    assert n_r == n_o and np.array_equal(a_r, a_o)


def _triangulation_case(seed):
    w, h = 640, 400
    left, right, _ = synth.stereo_pair(h, w, seed, d_min=5, d_max=30)
    e1, e2 = orbref.Extractor(1200), orbref.Extractor(1200)
    _, k1, d1 = e1(left)
    _, k2, d2 = e2(right)
    rng = np.random.default_rng(seed)
    views_ = []
    for k, d in ((k1, d1), (k2, d2)):
        node_of = (d[:, 0].astype(np.int64) >> 4) * 16 + (d[:, 1].astype(np.int64) >> 4)
        ids, inv = np.unique(node_of, return_inverse=True)
        order = np.argsort(inv, kind="stable")
        off = np.zeros(len(ids) + 1, np.int32)
        off[1:] = np.cumsum(np.bincount(inv, minlength=len(ids)))
        ur = np.where(rng.random(len(k)) < 0.5, k["x"] - 10, -1).astype(f32)
        hm = (rng.random(len(k)) < 0.2).astype(np.uint8)
        views_.append(orbref.make_keyframe_view(k, d, ur, hm, ids.astype(np.uint32), off, order.astype(np.uint32),
                                                e1.scale, e1.sigma2))
    return views_


@pytest.mark.parametrize("only_stereo,coarse,check,ep", [(False, False, True, (1e6, 200.0)), (True, False, True, (1e6, 200.0)),
                                                         (False, True, False, (320.0, 200.0)),
                                                         (False, False, False, (100.0, 150.0))])
def test_search_for_triangulation(only_stereo, coarse, check, ep, F12=None):
    """SearchForTriangulation, :886-1106 — row a15 (epipole gate, epipolar test, best-per-idx1, rotation histogram)."""
    v1, v2 = _triangulation_case(5)
    if F12 is None:
        F12 = np.array([[1e-7, 2e-6, -3e-4], [-2e-6, 1e-7, -1], [4e-4, 1, 2e-2]], f32)
    n_o, m_o = orbref.search_for_triangulation(v1, v2, F12, ep, only_stereo, coarse, check)
    n_r, m_r = refsrc.search_for_triangulation(v1, v2, F12, ep, only_stereo, coarse, check)
    assert n_o > 20
    assert n_r == n_o and np.array_equal(m_r, m_o)


_F12X4 = np.array([[[1e-7, 2e-6, -3e-4], [-2e-6, 1e-7, -1], [4e-4, 1, 2e-2]],          # left  x left
                   [[2e-7, 1e-6, -2e-4], [-1e-6, 2e-7, -1], [3e-4, 1, -3.0]],          # left  x right
                   [[1e-7, -2e-6, 3e-4], [2e-6, 1e-7, -1], [-4e-4, 1, 2.5]],           # right x left
                   [[3e-7, 2e-6, -1e-4], [-2e-6, 3e-7, -1], [1e-4, 1, -1.5]]], f32)     # right x right


@pytest.mark.parametrize("only_stereo,coarse,check,fl1,fl2", [(False, False, True, 0.5, 0.6), (False, False, False, 0.3, 1.0),
                                                              (False, True, True, 0.5, 0.5), (True, False, True, 0.5, 0.5),
                                                              (False, False, True, 0.0, 0.4)])
def test_search_for_triangulation_two_camera_keyframes(only_stereo, coarse, check, fl1, fl2):
    """SearchForTriangulation with both mpCamera2 set (:958-960, :966-971, :991-994, :996, :1007-1043): nothing is
    stereo, no epipole gate, the camera pair of the epipolar test follows the two features' sides of NLeft."""
    v1, v2 = _triangulation_case(5)
    nl1, nl2 = int(v1.struct.n * fl1), int(v2.struct.n * fl2)
    n_o, m_o = orbref.search_for_triangulation_fisheye(v1, nl1, v2, nl2, _F12X4, only_stereo, coarse, check)
    n_r, m_r = refsrc.search_for_triangulation_fisheye(v1, nl1, v2, nl2, _F12X4, only_stereo, coarse, check)
    assert n_r == n_o and np.array_equal(m_r, m_o)
    assert n_o == 0 if only_stereo else n_o > 20
    if not only_stereo and not coarse and 0 < fl1 < 1 and fl2 < 1:
        # the pair selection matters: one matrix for all four pairs gives a different answer
        n_x, m_x = orbref.search_for_triangulation_fisheye(v1, nl1, v2, nl2, np.stack([_F12X4[0]] * 4), False, False, check)
        assert not np.array_equal(m_x, m_o)


def test_triangulation_candidates_replay_equals_the_search():
    """orbref_triangulation_candidates + the host replay the drop-in body runs (bestDist = TH_LOW; a candidate at or
    below the running best that passes the epipolar test takes over, :988-1052) == the one-piece search."""
    v1, v2 = _triangulation_case(6)
    off, idx2, dist = orbref.triangulation_candidates(v1, v2)
    assert off[-1] == len(idx2) and len(idx2) > 100 and (dist <= 50).all()
    n_o, m_o = orbref.search_for_triangulation_fisheye(v1, 0, v2, 0, np.stack([_F12X4[3]] * 4), False, True, False)
    m = np.full(v1.struct.n, -1, np.int32)
    for i in range(v1.struct.n):
        best = 50
        for c in range(off[i], off[i + 1]):
            if dist[c] <= best:      # coarse: every candidate passes
                best, m[i] = dist[c], idx2[c]
    assert np.array_equal(m, m_o) and n_o == (m >= 0).sum()


@pytest.mark.parametrize("seed,nnratio,check", [(1, 0.7, True), (2, 0.9, False), (3, 0.6, True)])
def test_search_by_bow_both_overloads(seed, nnratio, check):
    """SearchByBoW(KeyFrame*, Frame&, ...) :230-404 and SearchByBoW(KeyFrame*, KeyFrame*, ...) :766-884."""
    k1, k2 = _keyframes(seed)
    n_o, m_o = orbref.search_by_bow(_view(k1), _view(k2), nnratio, check)
    n_r, m_r = refsrc.search_by_bow(_view(k1), _view(k2), nnratio, check)
    assert n_o > 20 and n_r == n_o and np.array_equal(m_r, m_o)
    n_o, m_o = orbref.search_by_bow_kf(_view(k1), _view(k2), nnratio, check)
    n_r, m_r = refsrc.search_by_bow_kf(_view(k1), _view(k2), nnratio, check)
    assert n_o > 20 and n_r == n_o and np.array_equal(m_r, m_o)


@pytest.mark.parametrize("seed,nnratio,check,kf_two", [(1, 0.7, True, True), (2, 0.9, False, False), (3, 0.6, True, True),
                                                       (4, 0.75, True, False)])
def test_search_by_bow_two_camera_frame(seed, nnratio, check, kf_two):
    """SearchByBoW(KeyFrame*, Frame&, ...) with F.Nleft != -1 (:274-365): left / right bests kept apart, the right one
    accepted without a ratio test inside the left one's TH_LOW gate; keypoints of rows >= NLeft from mvKeysRight."""
    k1, k2 = _keyframes(seed)
    k2, nl_f = _two_cameras(k2, seed)
    nl_kf = len(k1["kps"])
    if kf_two:
        k1, nl_kf = _two_cameras(k1, seed + 50)
    v1, v2 = _view(k1), _view(k2)
    n_o, m_o = orbref.search_by_bow_fisheye(v1, v2, nl_f, nnratio, check)
    n_r, m_r = refsrc.search_by_bow_fisheye(v1, nl_kf, v2, nl_f, nnratio, check)
    assert n_r == n_o and np.array_equal(m_r, m_o)
    assert (m_o[:nl_f] >= 0).sum() > 20 and (m_o[nl_f:] >= 0).sum() > 20
    # degenerate splits: everything left (== the one-camera function), everything right (no acceptance possible)
    n_a, m_a = orbref.search_by_bow_fisheye(v1, v2, v2.struct.n, nnratio, check)
    n_b, m_b = orbref.search_by_bow(v1, v2, nnratio, check)
    assert n_a == n_b and np.array_equal(m_a, m_b)
    n_c, m_c = refsrc.search_by_bow_fisheye(v1, nl_kf, v2, 0, nnratio, check)
    n_d, m_d = orbref.search_by_bow_fisheye(v1, v2, 0, nnratio, check)
    assert n_c == n_d == 0 and np.array_equal(m_c, m_d)


@pytest.mark.parametrize("seed,nnratio,check", [(1, 0.7, True), (2, 0.9, False)])
def test_search_by_bow_two_camera_keyframes(seed, nnratio, check):
    """SearchByBoW(KeyFrame*, KeyFrame*, ...) with NLeft != -1 (:799-801, :816-818): rows past mvKeysUn are skipped on
    both sides — the one-camera function on views whose MapPoint flags are cleared there."""
    k1, k2 = _keyframes(seed)
    un1, un2 = int(len(k1["kps"]) * 0.7), int(len(k2["kps"]) * 0.6)
    n_r, m_r = refsrc.search_by_bow_kf_fisheye(_view(k1), un1, _view(k2), un2, nnratio, check)
    c1, c2 = dict(k1), dict(k2)
    c1["hm"], c2["hm"] = k1["hm"].copy(), k2["hm"].copy()
    c1["hm"][un1:] = 0
    c2["hm"][un2:] = 0
    n_o, m_o = orbref.search_by_bow_kf(_view(c1), _view(c2), nnratio, check)
    assert n_o > 10 and n_r == n_o and np.array_equal(m_r, m_o)


def _fisheye_stereo_python(kl, dl, kr, dr, mono_l, mono_r, s2, R, t):
    """Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1271-1331) once more in numpy over the oracle's knn2, with the
    pseudo-triangulation of the stand-in KannalaBrandt8 (oracle/ref_stubs/matcher_world.h) in float32."""
    nl, nr = len(kl), len(kr)
    l2r, r2l = np.full(nl, -1, np.int32), np.full(nr, -1, np.int32)
    depth, p3d = np.full(nl, -1, f32), np.zeros((nl, 3), f32)
    i1, d1, i2, d2 = orbref.knn2(dl[mono_l:], dr[mono_r:])
    n = 0
    for q in range(nl - mono_l):
        if i2[q] < 0 or not (float(f32(d1[q])) < float(f32(d2[q])) * 0.7):
            continue
        a, b = q + mono_l, int(i1[q]) + mono_r
        d = f32(f32(f32(kl["x"][a] - kr["x"][b]) + f32(f32(0.25) * f32(kl["y"][a] - kr["y"][b]))) + f32(t[0]))
        z = f32(np.fmod(np.abs(d), f32(3.0)) - f32(1.0))
        if z > f32(0.0001):
            l2r[a], r2l[b], depth[a] = b, a, z
            p3d[a] = (f32(kl["x"][a] * s2[kl["octave"][a]]), f32(kr["y"][b] * s2[kr["octave"][b]]), f32(d * f32(R[0])))
            n += 1
    return n, l2r, r2l, depth, p3d


@pytest.mark.parametrize("seed,mono_l,mono_r", [(1, 0, 0), (2, 150, 90), (3, 399, 0), (4, 0, 398)])
def test_compute_stereo_fisheye_matches(seed, mono_l, mono_r):
    """Frame::ComputeStereoFishEyeMatches, src/Frame.cc:1271-1331 — the consumer of the brute-force 2-NN: Lowe's 0.7
    ratio on float distances (product in double), the triangulation call with both level sigmas, the bookkeeping."""
    rng = np.random.default_rng(seed)
    n = 400
    dl = synth.descriptors(n, seed)
    perm = rng.permutation(n)
    dr = synth.flip_bits(dl[perm], rng.integers(0, 60, n), rng)
    dr[:40] = dr[40:80]                       # duplicates: ratio-test ties (d1 == d2)
    kl, kr = np.zeros(n, synth.KP_DTYPE), np.zeros(n, synth.KP_DTYPE)
    for k in (kl, kr):
        k["x"], k["y"] = rng.uniform(0, 640, n).astype(f32), rng.uniform(0, 480, n).astype(f32)
        k["octave"] = rng.integers(0, 8, n)
    s2 = (f32(1.2) ** np.arange(8, dtype=f32)) ** 2
    R = np.array([0.75, 0, 0, 0, 1, 0, 0, 0, 1], f32)
    t = np.array([0.125, 0, 0], f32)
    got = refsrc.stereo_fisheye(kl, dl, kr, dr, mono_l, mono_r, s2, R, t)
    want = _fisheye_stereo_python(kl, dl, kr, dr, mono_l, mono_r, s2, R, t)
    assert got[0] == want[0]
    for g, w_ in zip(got[1:], want[1:]):
        assert np.array_equal(g, w_)
    if mono_l < 300 and mono_r < 300:
        assert got[0] > 30 and (got[3] < 0).sum() > 30


@pytest.mark.parametrize("seed,window,nnratio,check", [(6, 30, 0.9, True), (7, 60, 0.9, False), (8, 100, 0.7, True)])
def test_search_for_initialization(seed, window, nnratio, check):
    """SearchForInitialization, :618-764, with the serial executor in place of its tbb::parallel_for."""
    rng = np.random.default_rng(seed)
    n, w, h = 500, 640, 480
    k1, _, sf, _ = _frame(n, w, h, seed, False, 0.0)
    k1["octave"] = (rng.random(n) < 0.3).astype(np.int32) * rng.integers(1, 8, n)
    d1 = synth.flip_bits(synth.descriptors(n, seed, 40), rng.integers(0, 12, n), rng)
    perm = rng.permutation(n)
    k2 = k1[perm].copy()
    k2["x"] += rng.normal(0, 3, n).astype(f32)
    k2["y"] += rng.normal(0, 3, n).astype(f32)
    k2["angle"] = ((k2["angle"] + rng.normal(0, 20, n)) % 360).astype(f32)
    d2 = synth.flip_bits(d1[perm], rng.integers(0, 30, n), rng)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
    fvs = []
    for k, d in ((k1, d1), (k2, d2)):
        off, items = orbref.build_grid(k, 0.0, 0.0, inv_w, inv_h)
        g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
        fvs.append(orbref.make_frame_view(k, d, None, np.zeros(n, np.uint8), g, keep, sf))
    prev = np.stack([k1["x"], k1["y"]], axis=1)
    n_o, m_o = orbref.search_for_initialization(fvs[0], fvs[1], prev, window, nnratio, check)
    n_r, m_r = refsrc.search_for_initialization(fvs[0], fvs[1], prev, window, nnratio, check)
    assert n_o > 30 and n_r == n_o and np.array_equal(m_r, m_o)


def _projected_case(seed, m, th, stereo):
    kps, desc, sf, fv = _frame(700, 640, 480, seed, stereo)
    rng = np.random.default_rng(seed + 99)
    n = len(kps)
    src = rng.integers(0, n, m)
    anchored = rng.random(m) < 0.8
    u = np.where(anchored, kps["x"][src] + rng.normal(0, 2.0, m), rng.uniform(0, 640, m)).astype(f32)
    v = np.where(anchored, kps["y"][src] + rng.normal(0, 2.0, m), rng.uniform(0, 480, m)).astype(f32)
    # inside the image bounds: the reference tests uv against mnMinX.. itself (the caller-side filter of the projected form)
    u, v = np.clip(u, 0.5, 639.5).astype(f32), np.clip(v, 0.5, 479.5).astype(f32)
    octave = np.where(anchored, kps["octave"][src], rng.integers(0, 8, m)).astype(np.int32)
    angle = (np.where(anchored, kps["angle"][src] + rng.normal(0, 8.0, m), rng.uniform(0, 360, m)) % 360.0).astype(f32)
    d = synth.flip_bits(desc[src], rng.integers(0, 70, m), rng)
    return fv, sf, u, v, octave, angle, d, rng


@pytest.mark.parametrize("mode,stereo,th,check,seed", [(1, True, 7.0, True, 0), (-1, True, 7.0, True, 1),
                                                       (0, True, 15.0, True, 2), (0, False, 15.0, False, 3)])
def test_search_by_projection_last_frame(mode, stereo, th, check, seed):
    """SearchByProjection(Frame&, const Frame&, th, bMono), :1594-1806 — row a14: forward / backward / neither level
    windows, the uRight gate computed inside (uv(0) - mbf * invzc), already-matched rule, rotation histogram."""
    m, mbf, mb = 900, f32(47.9), f32(0.11)
    fv, sf, u, v, octave, angle, d, rng = _projected_case(seed, m, th, stereo)
    z = rng.uniform(0.5, 20.0, m).astype(f32)
    has_obs = (rng.random(m) < 0.9).astype(np.uint8)
    invz = (1.0 / z.astype(np.float64)).astype(f32)                       # const float invzc = 1.0 / x3Dc(2)
    ur = (u - (mbf * invz).astype(f32)).astype(f32) if stereo else None   # uv(0) - CurrentFrame.mbf * invzc
    lo = octave if mode > 0 else (np.zeros(m, np.int32) if mode < 0 else octave - 1)
    hi = np.full(m, -1, np.int32) if mode > 0 else (octave if mode < 0 else octave + 1)
    pts = orbref.make_projected(u, v, ur, (f32(th) * sf[octave]).astype(f32), lo.astype(np.int32), hi.astype(np.int32),
                                angle, has_obs, d)
    n_o, a_o = orbref.search_by_projection_frame(fv, pts, 100, check)
    n_r, a_r = refsrc.search_by_projection_last_frame(fv, u, v, z, octave, angle, has_obs, d, th, mbf, mb, mode, check)
    assert n_o > 30
    assert n_r == n_o and np.array_equal(a_r, a_o)


def _three_maxima(sizes):
    """ORBmatcher::ComputeThreeMaxima (:1920-1955) on the bin sizes."""
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, s in enumerate(sizes):
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if max2 < f32(0.1) * f32(max1):
        ind2 = ind3 = -1
    elif max3 < f32(0.1) * f32(max1):
        ind3 = -1
    return ind1, ind2, ind3


def _rot_bin(a_last, a_cur):
    rot = f32(a_last) - f32(a_cur)            # float rot = kpLF.angle - kpCF.angle
    if rot < 0.0:
        rot = f32(rot + f32(360.0))
    x = float(f32(rot * f32(1.0 / 30)))       # round(rot * factor): half away from zero
    b = int(np.floor(x + 0.5)) if x >= 0 else int(np.ceil(x - 0.5))
    return 0 if b == 30 else b


@pytest.mark.parametrize("mode,th,check,seed", [(0, 7.0, True, 20), (1, 7.0, True, 21), (-1, 10.0, True, 22),
                                                (0, 15.0, False, 23)])
def test_search_by_projection_last_frame_two_camera(mode, th, check, seed):
    """SearchByProjection(CurrentFrame, LastFrame, th, bMono) with CurrentFrame.Nleft != -1 (:1594-1806 incl. the
    right-camera block :1708-1780): per point the left search, then — unless the left window was empty (:1655) — the
    right search on mvKeysRight / mGridRight; the cameras write disjoint rows and share nmatches and ONE rotation
    histogram. The yardstick composes the oracle's one-camera loop (search_by_projection_frame_decisions) the way
    shim/ORBmatcher_orbx.cc composes the device's."""
    rng = np.random.default_rng(seed)
    nl, nr, m, w, h = 520, 480, 800, 640, 480
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
    kl = np.zeros(nl, synth.KP_DTYPE)
    kl["x"], kl["y"] = rng.uniform(0, w, nl).astype(f32), rng.uniform(0, h, nl).astype(f32)
    kl["octave"], kl["angle"] = rng.integers(0, 8, nl), rng.uniform(0, 360, nl).astype(f32)
    # the right camera sees most of the left features a few pixels away (GetRelativePoseTrl is the identity here)
    pick = rng.permutation(nl)[:nr]
    kr = kl[pick].copy()
    kr["x"] = np.clip(kr["x"] + rng.normal(0, 2.0, nr), 0, w - 1).astype(f32)
    kr["y"] = np.clip(kr["y"] + rng.normal(0, 2.0, nr), 0, h - 1).astype(f32)
    kr["angle"] = ((kr["angle"] + rng.normal(0, 4.0, nr)) % 360).astype(f32)
    dl = synth.descriptors(nl, seed)
    dr = synth.flip_bits(dl[pick], rng.integers(0, 25, nr), rng)
    occ = (rng.random(nl + nr) < 0.1).astype(np.uint8)
    fv = orbref.make_fisheye_view(kl, kr, np.concatenate([dl, dr]), occ, 0.0, 0.0, inv_w, inv_h,
                                  np.full(nl, -1, np.int32), np.full(nr, -1, np.int32), sf)
    src = rng.integers(0, nl, m)
    # a fifth of the points far from every feature: empty left windows, whose right search the reference skips
    far = rng.random(m) < 0.2
    u = np.where(far, rng.uniform(0, w, m), np.clip(kl["x"][src] + rng.normal(0, 2.0, m), 0.5, w - 0.5)).astype(f32)
    v = np.where(far, rng.uniform(0, h, m), np.clip(kl["y"][src] + rng.normal(0, 2.0, m), 0.5, h - 0.5)).astype(f32)
    z = rng.uniform(0.5, 20.0, m).astype(f32)
    octave = np.clip(kl["octave"][src] + rng.integers(-1, 2, m), 0, 7).astype(np.int32)
    angle = ((kl["angle"][src] + np.where(rng.random(m) < 0.8, rng.normal(0, 5.0, m), rng.uniform(0, 360, m))) % 360).astype(f32)
    has_obs = (rng.random(m) < 0.9).astype(np.uint8)
    d = synth.flip_bits(dl[src], rng.integers(0, 70, m), rng)
    mbf, mb = f32(47.9), f32(0.11)
    lo = octave if mode > 0 else (np.zeros(m, np.int32) if mode < 0 else octave - 1)
    hi = np.full(m, -1, np.int32) if mode > 0 else (octave if mode < 0 else octave + 1)
    radius = (f32(th) * sf[octave]).astype(f32)

    def view(k, dd, o):
        off, items = orbref.build_grid(k, 0.0, 0.0, inv_w, inv_h)
        g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
        return orbref.make_frame_view(k, dd, None, o, g, keep, sf)

    ptsL = orbref.make_projected(u, v, None, radius, lo.astype(np.int32), hi.astype(np.int32), angle, has_obs, d)
    decL, winL = orbref.search_by_projection_frame_decisions(view(kl, dl, occ[:nl]), ptsL, 100)
    keep = np.flatnonzero(winL > 0)
    assert 0 < len(keep) < m
    ptsR = orbref.make_projected(u[keep], v[keep], None, radius[keep], lo[keep].astype(np.int32),
                                 hi[keep].astype(np.int32), angle[keep], has_obs[keep], d[keep])
    decR_k, _ = orbref.search_by_projection_frame_decisions(view(kr, dr, occ[nl:]), ptsR, 100)
    decR = np.full(m, -1, np.int32)
    decR[keep] = decR_k
    want = np.full(nl + nr, -1, np.int32)
    hist = [[] for _ in range(30)]
    n_o = 0
    for i in range(m):
        if decL[i] >= 0:
            want[decL[i]] = i
            n_o += 1
            if check:
                hist[_rot_bin(angle[i], kl["angle"][decL[i]])].append(decL[i])
        if decR[i] >= 0:
            want[nl + decR[i]] = i
            n_o += 1
            if check:
                hist[_rot_bin(angle[i], kr["angle"][decR[i]])].append(nl + decR[i])
    if check:
        keep_bins = _three_maxima([len(b) for b in hist])
        for b in range(30):
            if b not in keep_bins:
                for row in hist[b]:
                    want[row] = -1
                    n_o -= 1
    n_r, a_r = refsrc.search_by_projection_last_frame_fisheye(fv, u, v, z, octave, angle, has_obs, d, th, mbf, mb, mode, check)
    assert n_o > 60 and (decR >= 0).sum() > 20
    assert n_r == n_o and np.array_equal(a_r, want)


@pytest.mark.parametrize("th,orb_dist,check,seed", [(10.0, 100, True, 4), (3.0, 64, True, 5), (10.0, 100, False, 6)])
def test_search_by_projection_keyframe(th, orb_dist, check, seed):
    """SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist), :1808-1918 (relocalisation): window
    [L-1, L+1], any MapPoint on the keypoint blocks, ORBdist, points in sAlreadyFound are skipped."""
    m = 900
    fv, sf, u, v, level, angle, d, rng = _projected_case(seed, m, th, False)
    found = (rng.random(m) < 0.15).astype(np.uint8)
    keep = np.flatnonzero(found == 0)
    pts = orbref.make_projected(u[keep], v[keep], None, (f32(th) * sf[level[keep]]).astype(f32),
                                (level[keep] - 1).astype(np.int32), (level[keep] + 1).astype(np.int32), angle[keep],
                                np.ones(len(keep), np.uint8), d[keep])
    n_o, a_o = orbref.search_by_projection_frame(fv, pts, orb_dist, check)
    n_r, a_r = refsrc.search_by_projection_keyframe(fv, u, v, level, angle, found, d, th, orb_dist, check)
    assert n_o > 30
    assert n_r == n_o and np.array_equal(a_r, np.where(a_o >= 0, keep[np.maximum(a_o, 0)], -1))


@pytest.mark.parametrize("sim3,th,seed", [(False, 3.0, 4), (True, 3.0, 5), (False, 2.5, 6), (True, 4.0, 7)])
def test_fuse_both_overloads(sim3, th, seed):
    """Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight) :1108-1281 (chi-square gate 7.8 / 5.99, level window
    [L-1, L], TH_LOW) and Fuse(KeyFrame*, Sophus::Sim3f&, ...) :1283-1390 (no gate): the keypoint each point is fused
    into, read off MapPoint::AddObservation, against the oracle's best_idx / best_dist."""
    rng = np.random.default_rng(seed)
    n, m, w, h = 600, 900, 640, 480
    kps = np.zeros(n, synth.KP_DTYPE)
    kps["x"], kps["y"] = rng.uniform(0, w, n).astype(f32), rng.uniform(0, h, n).astype(f32)
    kps["octave"] = rng.integers(0, 8, n)
    desc = synth.descriptors(n, seed)
    mbf = f32(47.9)
    ur_k = np.where(rng.random(n) < 0.5, kps["x"] - rng.uniform(1, 30, n), -1).astype(f32)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    kfv = orbref.make_frame_view(kps, desc, ur_k, np.zeros(n, np.uint8), g, keep, sf)
    src = rng.integers(0, n, m)
    u = np.clip(kps["x"][src] + rng.normal(0, 1.5, m), 0.5, w - 0.5).astype(f32)   # IsInImage is the caller-side filter
    v = np.clip(kps["y"][src] + rng.normal(0, 1.5, m), 0.5, h - 0.5).astype(f32)
    # depth consistent with the source keypoint's disparity where it has one, so that the stereo gate lets some through
    disp = np.where(ur_k[src] >= 0, kps["x"][src] - ur_k[src] + rng.normal(0, 1.0, m), rng.uniform(1, 30, m))
    z = (mbf / np.maximum(disp, 0.5)).astype(f32)
    lev = np.clip(kps["octave"][src] + rng.integers(-1, 2, m), 0, 7).astype(np.int32)
    d = synth.flip_bits(desc[src], rng.integers(0, 60, m), rng)
    invz = (f32(1) / z).astype(f32)                       # const float invz = 1 / p3Dc(2)
    pur = (u - (mbf * invz).astype(f32)).astype(f32)      # const float ur = uv(0) - bf * invz
    pts = orbref.make_projected(u, v, pur, (f32(th) * sf[lev]).astype(f32), lev - 1, lev, np.zeros(m, f32),
                                np.zeros(m, np.uint8), d)
    inv_s2 = (1.0 / (sf * sf)).astype(f32)
    bi, bd = orbref.fuse_match(kfv, inv_s2, pts, not sim3)
    n_r, best_r = refsrc.fuse(kfv, inv_s2, u, v, z, lev, d, th, mbf, sim3)
    want = np.where(bd <= 50, bi, -1)
    assert (want >= 0).sum() > 50
    assert n_r == (want >= 0).sum() and np.array_equal(best_r, want)


@pytest.mark.parametrize("b_right,th,seed", [(False, 3.0, 14), (True, 3.0, 15), (True, 2.5, 16)])
def test_fuse_two_camera_keyframe(b_right, th, seed):
    """Fuse(pKF, vpMapPoints, th, bRight) on a two-camera KeyFrame (NLeft != -1; :1116-1124, :1200-1201, :1219-1221,
    :1247): the search runs on ONE camera's keypoints, grid and descriptor rows — the left ones, or with bRight the
    right ones with the fused row offset by NLeft — so it is the oracle's fuse_match on that camera's view."""
    rng = np.random.default_rng(seed)
    nl, nr, m, w, h = 500, 430, 800, 640, 480
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)

    def cam(n, s):
        k = np.zeros(n, synth.KP_DTYPE)
        k["x"], k["y"] = rng.uniform(0, w, n).astype(f32), rng.uniform(0, h, n).astype(f32)
        k["octave"] = rng.integers(0, 8, n)
        return k, synth.descriptors(n, s)

    (kl, dl), (kr, dr) = cam(nl, seed), cam(nr, seed + 100)
    desc = np.concatenate([dl, dr])
    fv = orbref.make_fisheye_view(kl, kr, desc, np.zeros(nl + nr, np.uint8), 0.0, 0.0, inv_w, inv_h,
                                  np.full(nl, -1, np.int32), np.full(nr, -1, np.int32), sf)
    kc, dc, base = (kr, dr, nl) if b_right else (kl, dl, 0)
    src = rng.integers(0, len(kc), m)
    u = np.clip(kc["x"][src] + rng.normal(0, 1.5, m), 0.5, w - 0.5).astype(f32)
    v = np.clip(kc["y"][src] + rng.normal(0, 1.5, m), 0.5, h - 0.5).astype(f32)
    z = rng.uniform(1.0, 20.0, m).astype(f32)
    lev = np.clip(kc["octave"][src] + rng.integers(-1, 2, m), 0, 7).astype(np.int32)
    d = synth.flip_bits(dc[src], rng.integers(0, 60, m), rng)
    mbf = f32(47.9)
    pur = (u - (mbf * (f32(1) / z).astype(f32)).astype(f32)).astype(f32)
    inv_s2 = (1.0 / (sf * sf)).astype(f32)
    # the oracle on the searched camera's own view: mvuRight is all -1 on a two-camera KeyFrame (5.99 gate only)
    off, items = orbref.build_grid(kc, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    view = orbref.make_frame_view(kc, dc, np.full(len(kc), -1, f32), np.zeros(len(kc), np.uint8), g, keep, sf)
    pts = orbref.make_projected(u, v, pur, (f32(th) * sf[lev]).astype(f32), lev - 1, lev, np.zeros(m, f32),
                                np.zeros(m, np.uint8), d)
    bi, bd = orbref.fuse_match(view, inv_s2, pts, True)
    want = np.where(bd <= 50, bi + base, -1)
    n_r, best_r = refsrc.fuse_two_camera(fv, inv_s2, u, v, z, lev, d, th, mbf, b_right)
    assert (want >= 0).sum() > 50
    assert n_r == (want >= 0).sum() and np.array_equal(best_r, want)


@pytest.mark.parametrize("with_kfs,th,ratio,seed", [(False, 8, 1.0, 8), (True, 8, 1.0, 9), (False, 4, 1.5, 10)])
def test_sim3_search_by_projection_is_the_projected_form(with_kfs, th, ratio, seed):
    """SearchByProjection(KeyFrame*, Sim3f&, vpPoints, vpMatched, th, ratioHamming) :406-506 and its :508-616 twin are
    the projected search with window [L-1, L], max_dist = TH_LOW * ratioHamming, no rotation check, every already
    matched keypoint closed and every written keypoint closing (the mapping INTEGRATION.md gives for them)."""
    m = 900
    fv, sf, u, v, level, angle, d, rng = _projected_case(seed, m, float(th), False)
    kps_n = fv.struct.n
    matched_in = (rng.random(kps_n) < 0.1).astype(np.uint8)
    # (the reference ignores the frame view's own `occupied` here; the oracle gets matched_in in that role below)
    n_r, a_r = refsrc.search_by_projection_sim3(fv, matched_in, u, v, level, d, th, ratio, with_kfs)
    fv2 = _with_occupied(seed, matched_in)
    pts = orbref.make_projected(u, v, None, (f32(th) * sf[level]).astype(f32), (level - 1).astype(np.int32), level,
                                angle, np.ones(m, np.uint8), d)
    n_o, a_o = orbref.search_by_projection_frame(fv2, pts, int(np.floor(50 * ratio)), False)
    assert n_o > 30
    assert n_r == n_o and np.array_equal(a_r, a_o)


def _with_occupied(seed, occupied):
    """The frame of _projected_case(seed, ...) with another occupied mask."""
    kps, desc, sf, _ = _frame(700, 640, 480, seed, False)
    inv_w, inv_h = f32(64) / f32(640), f32(48) / f32(480)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    return orbref.make_frame_view(kps, desc, None, np.ascontiguousarray(occupied, np.uint8), g, keep, sf)


@pytest.mark.parametrize("th,seed", [(7.5, 11), (4.0, 12)])
def test_search_by_sim3_is_two_gate_free_fuse_matches_plus_agreement(th, seed):
    """SearchBySim3, :1392-1592 = per direction the gate-free matching loop of orbref.fuse_match with
    bestDist <= TH_HIGH, then the mutual-agreement pass of :1575-1588 (the mapping INTEGRATION.md gives for it)."""
    rng = np.random.default_rng(seed)
    n, w, h = 500, 640, 480
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
    k1 = np.zeros(n, synth.KP_DTYPE)
    k1["x"], k1["y"] = rng.uniform(0, w, n).astype(f32), rng.uniform(0, h, n).astype(f32)
    k1["octave"] = rng.integers(0, 8, n)
    d1 = synth.descriptors(n, seed)
    perm = rng.permutation(n)
    k2 = k1[perm].copy()
    k2["x"] += rng.normal(0, 1.5, n).astype(f32)
    k2["y"] += rng.normal(0, 1.5, n).astype(f32)
    d2 = synth.flip_bits(d1[perm], rng.integers(0, 60, n), rng)
    views_, sides = [], []
    for k, d in ((k1, d1), (k2, d2)):
        off, items = orbref.build_grid(k, 0.0, 0.0, inv_w, inv_h)
        g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
        views_.append(orbref.make_frame_view(k, d, None, np.zeros(n, np.uint8), g, keep, sf))
        has = (rng.random(n) < 0.8).astype(np.uint8)
        # where the feature's MapPoint projects in the OTHER KeyFrame (inside its image: IsInImage)
        u = np.clip(k["x"] + rng.normal(0, 1.0, n), 0.5, w - 0.5).astype(f32)
        v = np.clip(k["y"] + rng.normal(0, 1.0, n), 0.5, h - 0.5).astype(f32)
        level = np.clip(k["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
        sides.append((has, u, v, level, synth.flip_bits(d, rng.integers(0, 20, n), rng)))
    n_r, m_r = refsrc.search_by_sim3(views_[0], views_[1], sides[0], sides[1], th)

    def direction(src, dst_view):
        has, u, v, level, d = src
        keep = np.flatnonzero(has)
        pts = orbref.make_projected(u[keep], v[keep], None, (f32(th) * sf[level[keep]]).astype(f32),
                                    (level[keep] - 1).astype(np.int32), level[keep], np.zeros(len(keep), f32),
                                    np.zeros(len(keep), np.uint8), d[keep])
        bi, bd = orbref.fuse_match(dst_view, (1.0 / (sf * sf)).astype(f32), pts, False)
        out = np.full(n, -1, np.int32)
        out[keep] = np.where(bd <= 100, bi, -1)
        return out
    match1, match2 = direction(sides[0], views_[1]), direction(sides[1], views_[0])
    want = np.full(n, -1, np.int32)
    for i1 in range(n):
        if match1[i1] >= 0 and match2[match1[i1]] == i1:
            want[i1] = match1[i1]
    assert (want >= 0).sum() > 30
    assert n_r == (want >= 0).sum() and np.array_equal(m_r, want)


@pytest.mark.parametrize("w,h,nfeat,seed,kind", [(752, 480, 1200, 1, "scene"), (640, 480, 1000, 2, "scene"),
                                                 (752, 480, 1200, 3, "noise_blur"), (400, 300, 600, 8, "scene")])
def test_stereo_frame_hot_path_equals_the_reference_source(w, h, nfeat, seed, kind):
    """Row a12 and the headline path end to end: ORBextractor::operator() on both images + Frame::ComputeStereoMatches
    (src/Frame.cc:921-1084: row table, Hamming search, 11x11 SAD slide on the raw pyramid levels, parabola, median
    filter), every line of it the reference's own text, against the oracle's extract x2 + stereo_match."""
    left, right, _ = synth.stereo_pair(h, w, seed, kind=kind)
    mbf, mb = float(f32(435.2 * 0.11)), float(f32(0.11))
    n_r, kl_r, dl_r, kr_r, dr_r, ur_r, dp_r = refsrc.stereo_frame(left, right, mbf, mb, nfeat)
    ex_l, ex_r = orbref.Extractor(nfeat), orbref.Extractor(nfeat)
    _, kl, dl = ex_l(left)
    _, kr, dr = ex_r(right)
    assert np.array_equal(kl, kl_r) and np.array_equal(dl, dl_r) and np.array_equal(kr, kr_r) and np.array_equal(dr, dr_r)
    n_o, ur_o, dp_o = orbref.stereo_match(ex_l, ex_r, kl, dl, kr, dr, mbf, mb)
    assert n_o > 50
    assert n_r == n_o
    assert np.array_equal(ur_r.view(np.uint32), ur_o.view(np.uint32))
    assert np.array_equal(dp_r.view(np.uint32), dp_o.view(np.uint32))


def test_grid_functions():
    """Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:520-547, 833-844), Frame::GetFeaturesInArea (:765-831) and
    KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:705-749) — §8(f) rank 1 — against orbref.build_grid / features_in_area."""
    rng = np.random.default_rng(41)
    n, w, h = 900, 640, 480
    kps = np.zeros(n, synth.KP_DTYPE)
    kps["x"] = rng.uniform(-15, w + 15, n).astype(f32)       # undistorted keypoints may leave the image (:838-842)
    kps["y"] = rng.uniform(-15, h + 15, n).astype(f32)
    kps["x"][:40] = (np.round(kps["x"][:40] / 10) * 10 + 5).astype(f32)   # on cell-rounding boundaries (w / 64 = 10)
    kps["octave"] = rng.integers(0, 8, n)
    desc = synth.descriptors(n, 41)
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    fv = orbref.make_frame_view(kps, desc, None, np.zeros(n, np.uint8), g, keep, sf)
    off_r, items_r = refsrc.build_grid(fv)
    assert np.array_equal(off_r, off) and np.array_equal(items_r, items[:off[-1]])
    for _ in range(400):
        x, y = float(rng.uniform(-30, w + 30)), float(rng.uniform(-30, h + 30))
        r = float(rng.choice([2.5, 7.0, 15.0, 40.0, 100.0]))
        lo, hi = [(-1, -1), (0, 0), (3, -1), (0, 4), (2, 3), (5, 6)][int(rng.integers(0, 6))]
        want = orbref.features_in_area(fv, x, y, r, lo, hi)
        assert np.array_equal(refsrc.features_in_area(fv, x, y, r, lo, hi), want)
        assert np.array_equal(refsrc.features_in_area(fv, x, y, r, keyframe=True), orbref.features_in_area(fv, x, y, r, -1, -1))


def test_compute_distinctive_descriptors():
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:372-441): all-pairs distances, per row the element
    [0.5 (N - 1)] of the sorted row, first row with the least such median."""
    rng = np.random.default_rng(51)
    assert refsrc.distinctive_descriptor(np.zeros((0, 32), np.uint8)) is None and orbref.distinctive_descriptor(np.zeros((0, 32), np.uint8)) == -1
    for n in list(range(1, 12)) + [17, 30, 64, 101]:
        for proto in (0, 3):
            base = synth.descriptors(1, int(rng.integers(1 << 30)))[0]
            d = synth.flip_bits(np.repeat(base[None], n, 0), rng.integers(0, 60, n), rng)
            if proto:
                d[rng.integers(0, n, max(1, n // 3))] = d[0]       # duplicates: equal medians, the first row must win
            got = refsrc.distinctive_descriptor(d)
            idx = orbref.distinctive_descriptor(d)
            assert idx >= 0 and np.array_equal(got, d[idx]), (n, proto)


@pytest.mark.parametrize("seed,m,cos_limit", [(0, 10000, 0.5), (1, 10000, 0.5), (2, 3000, 0.8), (3, 1, 0.5), (4, 500, -1.0)])
def test_is_in_frustum(seed, m, cos_limit):
    """Frame::isInFrustum (src/Frame.cc:632-699) + MapPoint::PredictScale (src/MapPoint.cc:559-573): the reference's own
    text (cut out by signature, oracle/Makefile) on a general pose must leave exactly the oracle's words on every point:
    mbTrackInView, mTrackProjX / Y (-1 outside the image), and — only when in view — mTrackProjXR, mnTrackScaleLevel,
    mTrackViewCos, mTrackDepth. The stand-in Eigen reduces 3-vectors in Eigen's own order c0 + (c1 + c2)."""
    from orb_slam3_fast_b200 import synth
    img = synth.scene(480, 640, seed=20 + seed)
    _, kps, desc = orbref.Extractor(1200)(img, (0, 0))
    fr = synth.frustum(640, 480, seed=seed)
    mp = synth.local_map_world(kps, desc, m, fr, seed=seed)
    lm = orbref.make_local_map(**mp)
    sentinel = lambda: dict(track_in_view=np.full(m, 7, np.uint8), proj_x=np.full(m, 3.5, np.float32),
                            proj_y=np.full(m, 4.5, np.float32), proj_xr=np.full(m, 5.5, np.float32),
                            level=np.full(m, -9, np.int32), view_cos=np.full(m, 6.5, np.float32),
                            depth=np.full(m, 8.5, np.float32))
    nv_r, o_r = refsrc.is_in_frustum(fr, lm, 0, cos_limit, sentinel())
    nv, o = orbref.is_in_frustum(fr, lm, 0, cos_limit, sentinel())
    assert nv == nv_r and (m < 100 or cos_limit > 0.6 or nv > m // 3)
    for k in o:
        assert o[k].tobytes() == o_r[k].tobytes(), k
    # untouched fields really are untouched: points out of view keep the sentinel in the "only when in view" fields
    out = o["track_in_view"] == 0
    if out.any():
        assert (o["level"][out] == -9).all() and (o["depth"][out] == 8.5).all()


# (mSensor, isImuInitialized, GetIniertialBA2, mState, frame id, mnLastRelocFrameId, mbFarPoints) -> th of :3302-3322
_SLP_CASES = [((1, 0, 0, 2, 40, 0, 1), 1), ((2, 0, 0, 2, 40, 0, 0), 3), ((4, 1, 1, 2, 40, 0, 1), 2), ((4, 1, 0, 2, 40, 0, 0), 6),
              ((3, 0, 0, 2, 40, 0, 1), 10), ((1, 0, 0, 2, 40, 39, 0), 5), ((5, 1, 1, 3, 40, 39, 1), 15), ((0, 0, 0, 4, 7, 0, 0), 15)]


@pytest.mark.parametrize("seed,m,case", [(0, 10000, 0), (1, 4000, 1), (2, 4000, 2), (3, 3000, 3), (4, 3000, 4), (5, 3000, 5),
                                         (6, 2000, 6), (7, 2000, 7), (8, 1, 0), (9, 0, 0)])
def test_search_local_points(seed, m, case):
    """void Tracking::SearchLocalPoints() (src/Tracking.cc:3249-3330), the caller of a13: the
    reference's own function text on a stand-in Tracking object (oracle/Makefile pipes it in) against the oracle's pieces
    put together here — the bookkeeping loop over mCurrentFrame.mvpMapPoints (:3268-3285), isInFrustum over
    mvpLocalMapPoints with the skip rule (:3288-3300), the search radius of the tracker's state (:3302-3322) and
    SearchByProjection (:3324). Compared: every slot of mvpMapPoints, every tracking word of every MapPoint (incl. the ones
    the reference leaves untouched, and stale mbTrackInView flags on bad points), mnVisible, mnLastFrameSeen and
    mCurrentFrame.mmProjectPoints. The shim worlds run the same comparison on the drop-in body (shim/Tracking_orbx.cc)."""
    ctl, th = _SLP_CASES[case]
    rng = np.random.default_rng(900 + seed)
    img = synth.scene(480, 640, seed=30 + seed)
    ex = orbref.Extractor(1200)
    _, kps, desc = ex(img, (0, 0))
    n = len(kps)
    fr = synth.frustum(640, 480, seed=seed)
    mp = synth.local_map_world(kps, desc, max(m, 1), fr, seed=seed)
    if m == 0:
        mp = {k: v[:0] for k, v in mp.items()}
    frame_id = ctl[4]
    bad = (rng.random(m) < 0.04).astype(np.uint8)
    held = np.full(n, -1, np.int32)
    if m:
        slots = rng.choice(n, min(n // 8, m), replace=False)
        held[slots] = rng.choice(m, len(slots), replace=False)       # some of them bad, some without observations
    held[rng.choice(n, n // 20, replace=False)] = -2
    state = dict(track_in_view=(rng.random(m) < 0.3).astype(np.uint8), proj_x=np.full(m, 3.5, f32), proj_y=np.full(m, 4.5, f32),
                 proj_xr=np.full(m, 5.5, f32), level=np.full(m, -9, np.int32), view_cos=np.full(m, 6.5, f32),
                 depth=np.full(m, 8.5, f32), visible=rng.integers(0, 50, m).astype(np.int32),
                 last_seen=np.where(rng.random(m) < 0.05, frame_id, frame_id - 1).astype(np.int32))
    # mnLastFrameSeen == mnId is only ever written together with mbTrackInView = false (:3278-3280), so a live point seen
    # in this frame cannot carry a stale flag (a bad point can: nothing touches it any more)
    state["track_in_view"][(state["last_seen"] == frame_id) & (bad == 0)] = 0
    inv_w, inv_h = f32(64) / f32(640), f32(48) / f32(480)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    ur = np.where(rng.random(n) < 0.6, kps["x"] - rng.uniform(1, 40, n), -1).astype(f32)
    lm0 = orbref.make_local_map(**{**mp, "skip": None})
    fv0 = orbref.make_frame_view(kps, desc, ur, np.zeros(n, np.uint8), g, keep, ex.scale)
    a_r, s_r, pp_r, k_r = refsrc.search_local_points(fv0, fr, lm0, held, bad, ctl, 15.0, state)

    # ---- the same pass from the oracle's pieces ----
    st = {k: v.copy() for k, v in state.items()}
    now = held.copy()
    for i in np.flatnonzero(held >= 0):                            # :3268-3285
        k = held[i]
        if bad[k]:
            now[i] = -1
        else:
            st["visible"][k] += 1
            st["last_seen"][k] = frame_id
            st["track_in_view"][k] = 0
    skip = ((st["last_seen"] == frame_id) | (bad != 0)).astype(np.uint8)    # :3289
    pp = np.full((m, 2), np.nan, f32)
    nv = 0
    if m:
        lm = orbref.make_local_map(**{**mp, "skip": skip})
        before = st["track_in_view"].copy()
        nv, o = orbref.is_in_frustum(fr, lm, 0, 0.5, {k: st[k] for k in ("track_in_view", "proj_x", "proj_y", "proj_xr",
                                                                         "level", "view_cos", "depth")})
        for k in o:
            st[k] = o[k]
        st["track_in_view"] = np.where(skip != 0, before, st["track_in_view"])   # a skipped point is not touched at all
        seen = (o["track_in_view"] != 0) & (skip == 0)
        st["visible"][seen] += 1
        pp[seen, 0], pp[seen, 1] = st["proj_x"][seen], st["proj_y"][seen]
    want = now.copy()
    if nv > 0:
        occ = np.array([h == -2 or (h >= 0 and mp["has_obs"][h] != 0) for h in now], np.uint8)
        fv = orbref.make_frame_view(kps, desc, ur, occ, g, keep, ex.scale)
        mps = orbref.make_mappoints(st["track_in_view"] & (1 - bad), st["proj_x"], st["proj_y"], st["proj_xr"], st["level"],
                                    st["view_cos"], st["depth"], mp["has_obs"], mp["desc"])
        nm, assign = orbref.search_by_projection_map(fv, mps, float(th), 0.8, bool(ctl[6]), 15.0)
        want = np.where(assign >= 0, assign, now)
        assert m < 1000 or nm > 50, "the case must produce matches"
    assert np.array_equal(a_r, want)
    for k in st:
        assert s_r[k].tobytes() == st[k].tobytes(), k
    assert pp_r.tobytes() == pp.tobytes() and k_r == int(np.isfinite(pp[:, 0]).sum())
    if m >= 1000:
        assert nv > m // 4 and (skip != 0).sum() > 0 and (bad & state["track_in_view"]).sum() > 0


def test_search_local_points_two_camera_frame():
    """The Nleft != -1 side of Tracking::SearchLocalPoints: Frame::isInFrustum (src/Frame.cc:687-698) resets
    mbTrackInView / mnTrackScaleLevel of every point it is asked about and asks isInFrustumChecks, which in the stand-in
    world sees nothing — so what is observable is the bookkeeping: the loop over the frame's own points, the skip rule, and
    that skipped and bad points keep every word. (The drop-in body keeps the reference's per-point call on such frames.)"""
    rng = np.random.default_rng(77)
    img = synth.scene(480, 640, seed=41)
    ex = orbref.Extractor(1200)
    _, kps, desc = ex(img, (0, 0))
    n, m, frame_id = len(kps), 1500, 40
    fr = synth.frustum(640, 480, seed=3)
    mp = synth.local_map_world(kps, desc, m, fr, seed=3)
    bad = (rng.random(m) < 0.05).astype(np.uint8)
    held = np.full(n, -1, np.int32)
    slots = rng.choice(n, 100, replace=False)
    held[slots] = rng.choice(m, 100, replace=False)
    state = dict(track_in_view=(rng.random(m) < 0.3).astype(np.uint8), proj_x=np.full(m, 3.5, f32), proj_y=np.full(m, 4.5, f32),
                 proj_xr=np.full(m, 5.5, f32), level=np.full(m, 2, np.int32), view_cos=np.full(m, 6.5, f32),
                 depth=np.full(m, 8.5, f32), visible=rng.integers(0, 50, m).astype(np.int32),
                 last_seen=np.where(rng.random(m) < 0.05, frame_id, frame_id - 1).astype(np.int32))
    state["track_in_view"][(state["last_seen"] == frame_id) & (bad == 0)] = 0
    inv_w, inv_h = f32(64) / f32(640), f32(48) / f32(480)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    fv = orbref.make_frame_view(kps, desc, None, np.zeros(n, np.uint8), g, keep, ex.scale)
    lm = orbref.make_local_map(**{**mp, "skip": None})
    a_r, s_r, pp_r, k_r = refsrc.search_local_points(fv, fr, lm, held, bad, (1, 0, 0, 2, frame_id, 0, 0, 1), 15.0, state)
    st = {k: v.copy() for k, v in state.items()}
    now = held.copy()
    for i in np.flatnonzero(held >= 0):
        k = held[i]
        if bad[k]:
            now[i] = -1
        else:
            st["visible"][k] += 1
            st["last_seen"][k] = frame_id
            st["track_in_view"][k] = 0
    asked = (st["last_seen"] != frame_id) & (bad == 0)
    st["track_in_view"][asked] = 0
    st["level"][asked] = -1
    assert asked.sum() > m // 2 and (~asked).sum() > 100
    assert np.array_equal(a_r, now) and k_r == 0 and np.isnan(pp_r).all()
    for k in st:
        assert s_r[k].tobytes() == st[k].tobytes(), k


@pytest.mark.parametrize("seed,th,far,sizes", [(0, 1.0, True, (700, 650, 3000)), (1, 3.0, False, (700, 650, 3000)),
                                               (2, 6.0, True, (1200, 1100, 10000)), (4, 15.0, False, (300, 280, 2000))])
def test_search_by_projection_map_fisheye(seed, th, far, sizes):
    """SearchByProjection(Frame&, vector<MapPoint*>) with Nleft != -1 (:42-221 incl. the right-camera twin :148-217): the
    reference's own method on a two-camera stand-in Frame (grids from its own AssignFeaturesToGrid) against the oracle's
    restatement — every slot of mvpMapPoints and the returned count, with stereo partners, overwrites and slots that
    re-open when a point without observations replaces one with."""
    fr, mp, mpr = synth.fisheye_case(*sizes, seed=seed)
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    fv = orbref.make_fisheye_view(fr["kps_left"], fr["kps_right"], fr["desc"], fr["occupied"], 0.0, 0.0, f32(64) / f32(640),
                                  f32(48) / f32(480), fr["left_to_right"], fr["right_to_left"], sf)
    mps, mr = orbref.make_mappoints(**mp), orbref.make_mappoints_right(**mpr)
    n_o, a_o = orbref.search_by_projection_map_fisheye(fv, mps, mr, th, 0.8, far, 15.0)
    n_r, a_r = refsrc.search_by_projection_map_fisheye(fv, mps, mr, th, 0.8, far, 15.0)
    assert n_o > 300 and n_o > (a_o >= 0).sum(), "the case must contain partner writes"
    assert n_r == n_o and np.array_equal(a_r, a_o)
