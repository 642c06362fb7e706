"""Parity of the batched local-map tracking search (BASELINE.json configs[3] as a throughput path) with the CPU oracle,
through the C ABI: Frame::isInFrustum (src/Frame.cc:632-699), Frame::AssignFeaturesToGrid and
ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, ...) (src/ORBmatcher.cc:42-221) in serial MapPoint
order, for frames that never leave the device (orbm_track_local_map_batch_device) and end to end from host buffers
(orbm_stereo_track_frames_batch). Every output word — float bits included — must equal the oracle's."""
import ctypes as C

import numpy as np
import pytest

from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, synth, views
from orb_slam3_fast_b200.lib import OrbxError
from oracle import orbref

pytestmark = pytest.mark.gpu

W, H, NF = 640, 480, 1200
MBF, MB = float(np.float32(435.2 * 0.11)), float(np.float32(0.11))


@pytest.fixture(scope="module")
def matcher(gpu):
    return ORBmatcher(0.8, True)


def _same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


def test_is_in_frustum_equals_oracle(matcher):
    img = synth.scene(H, W, seed=3)
    _, kps, desc = orbref.Extractor(NF)(img, (0, 0))
    for seed, m in ((5, 10000), (6, 1), (7, 257)):
        fr = synth.frustum(W, H, seed=seed)
        mp = synth.local_map_world(kps, desc, m, fr, seed=seed)
        nv_r, o_r = orbref.is_in_frustum(fr, orbref.make_local_map(**mp))
        nv, o = matcher.IsInFrustum(fr, views.make_local_map(**mp))
        assert nv == nv_r and (m < 100 or nv_r > m // 3)
        for k in o_r:
            assert _same(o[k], o_r[k]), "%s differs at %s" % (k, np.nonzero(o[k] != o_r[k])[0][:8])
    # second map of a two-map set; no skip array; a viewing-angle limit nothing passes
    fr = synth.frustum(W, H, seed=9)
    a, b = synth.local_map_world(kps, desc, 500, fr, seed=1), synth.local_map_world(kps, desc, 500, fr, seed=2)
    two = {k: np.stack([a[k], b[k]]) for k in a}
    two["skip"] = None
    nv_r, o_r = orbref.is_in_frustum(fr, orbref.make_local_map(**two), 1)
    nv, o = matcher.IsInFrustum(fr, views.make_local_map(**two), 1)
    assert nv == nv_r > 100 and all(_same(o[k], o_r[k]) for k in o_r)
    nv, o = matcher.IsInFrustum(fr, views.make_local_map(**two), 0, viewingCosLimit=1.5)
    assert nv == 0 and not o["track_in_view"].any()
    nv, _ = matcher.IsInFrustum(fr, views.make_local_map(**{k: (None if v is None else v[:, :0]) for k, v in two.items()}))
    assert nv == 0


def _frames(n_frames, seed0):
    """n_frames seeded stereo pairs with their oracle keypoints / descriptors / stereo matches."""
    L, R, ref = [], [], []
    rl, rr = orbref.Extractor(NF), orbref.Extractor(NF)
    for s in range(n_frames):
        kind = "scene" if s % 3 else "noise_blur"
        left, right, _ = synth.stereo_pair(H, W, seed0 + s, kind=kind)
        _, kl, dl = rl(left, (0, 0))
        _, kr, dr = rr(right, (0, 0))
        nm, ur, dp = orbref.stereo_match(rl, rr, kl, dl, kr, dr, MBF, MB)
        L.append(left)
        R.append(right)
        ref.append(dict(kl=kl, dl=dl, kr=kr, dr=dr, nm=nm, ur=ur, dp=dp))
    return np.stack(L), np.stack(R), ref


def _oracle_track(ref, fr, mp_one, occ, th, nnratio, far, th_far):
    kl, dl = ref["kl"], ref["dl"]
    prm = views.make_track_params(W, H)
    off, items = orbref.build_grid(kl, 0.0, 0.0, prm.inv_w, prm.inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, prm.inv_w, prm.inv_h)
    fv = orbref.make_frame_view(kl, dl, ref["ur"], occ, g, keep, orbref.Extractor(NF).scale)
    return orbref.track_local_map(fv, fr, orbref.make_local_map(**mp_one), 0, th, nnratio, far, th_far)


@pytest.mark.parametrize("th,far,m", [(1.0, False, 10000), (3.0, True, 3000), (15.0, False, 600)])
def test_stereo_track_frames_batch_equals_oracle(matcher, th, far, m):
    """Host buffers in and out: extract x2 + ComputeStereoMatches + SearchLocalPoints, 7 pairs in groups of 3 (the
    pipeline rotates through its lanes), one local map per pair; pairs 5 and 6 share map 5 through map_index."""
    n = 7
    L, R, ref = _frames(n, 40)
    frs = np.stack([synth.frustum(W, H, seed=100 + i) for i in range(n)])
    n_maps = 6
    maps = [synth.local_map_world(ref[i]["kl"], ref[i]["dl"], m, frs[i], seed=200 + i) for i in range(n_maps)]
    stacked = {k: np.stack([mp[k] for mp in maps]) for k in maps[0]}
    map_index = np.array([0, 1, 2, 3, 4, 5, 5], np.int32)
    frs[6] = frs[5]  # the pose the shared map was generated for
    rng = np.random.default_rng(1)
    exl, exr = ORBextractor(NF, max_batch=3), ORBextractor(NF, max_batch=3)
    cap = exl.capacity
    occ = (rng.random((n, cap)) < 0.25).astype(np.uint8)
    prm = views.make_track_params(W, H, th=th, nnratio=0.8, far_points=far, th_far=12.0)
    out = matcher.StereoTrackFramesBatch(exl, exr, L, R, MBF, MB, frs, views.make_local_map(**stacked), prm,
                                         map_index=map_index, occupied=occ)
    total = 0
    for i in range(n):
        r = ref[i]
        nl = len(r["kl"])
        assert int(out["n_l"][i]) == nl and int(out["n_r"][i]) == len(r["kr"])
        assert _same(out["kps_l"][i, :nl], r["kl"]) and _same(out["desc_l"][i, :nl], r["dl"])
        assert int(out["n_matched"][i]) == r["nm"] and _same(out["u_right"][i, :nl], r["ur"])
        assert _same(out["depth"][i, :nl], r["dp"])
        nm, assign, nv = _oracle_track(r, frs[i], maps[map_index[i]], occ[i, :nl], th, 0.8, far, 12.0)
        assert int(out["n_in_view"][i]) == nv
        assert int(out["nmatches"][i]) == nm, "pair %d: nmatches %d vs %d" % (i, out["nmatches"][i], nm)
        assert np.array_equal(out["assign"][i, :nl], assign), "pair %d assign differs" % i
        total += nm
    assert total > 50 * n, "degenerate test: only %d matches" % total


def test_track_local_map_batch_device_equals_oracle(matcher):
    """The device-resident form on the outputs of orbx_extract_batch_device + orbm_stereo_match_batch_device; default
    map rule (frame f -> map f % n_maps), no occupancy array, monocular (no uRight) as a second pass."""
    import torch
    n = 5
    L, R, ref = _frames(n, 60)
    dev = torch.device("cuda", 0)
    exl, exr = ORBextractor(NF, max_batch=n), ORBextractor(NF, max_batch=n)
    cap = exl.capacity
    dL, dR = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)

    def outs():
        return dict(kps=torch.empty((n, cap, 7), dtype=torch.int32, device=dev),
                    desc=torch.empty((n, cap, 32), dtype=torch.uint8, device=dev),
                    n=torch.empty(n, dtype=torch.int32, device=dev), mono=torch.empty(n, dtype=torch.int32, device=dev),
                    status=torch.empty(n, dtype=torch.int32, device=dev))
    oL, oR = outs(), outs()
    st = torch.cuda.Stream(device=dev)
    s = st.cuda_stream
    for ex, imgs, o in ((exl, dL, oL), (exr, dR, oR)):
        ex.extract_batch_device(imgs.data_ptr(), n, W, H, W, W * H, (0, 0), o["kps"].data_ptr(), o["desc"].data_ptr(), cap,
                                o["n"].data_ptr(), o["mono"].data_ptr(), o["status"].data_ptr(), s)
    d_ur = torch.empty((n, cap), dtype=torch.float32, device=dev)
    d_dp = torch.empty((n, cap), dtype=torch.float32, device=dev)
    d_nm = torch.empty(n, dtype=torch.int32, device=dev)
    matcher.ComputeStereoMatches_device(exl, exr, n, oL["kps"].data_ptr(), oL["desc"].data_ptr(), oL["n"].data_ptr(),
                                        oR["kps"].data_ptr(), oR["desc"].data_ptr(), oR["n"].data_ptr(), cap, MBF, MB,
                                        d_ur.data_ptr(), d_dp.data_ptr(), d_nm.data_ptr(), s)
    m, n_maps = 4000, 5
    frs = np.stack([synth.frustum(W, H, seed=300 + i) for i in range(n)])
    maps = [synth.local_map_world(ref[i]["kl"], ref[i]["dl"], m, frs[i], seed=400 + i) for i in range(n_maps)]
    t = {k: torch.from_numpy(np.ascontiguousarray(np.stack([mp[k] for mp in maps]))).to(dev) for k in maps[0]}
    d_fr = torch.from_numpy(frs.view(np.uint8).reshape(n, -1)).to(dev)
    dmap = views.make_local_map_device(m, n_maps, t["pos"].data_ptr(), t["normal"].data_ptr(), t["min_dist"].data_ptr(),
                                       t["max_dist"].data_ptr(), t["skip"].data_ptr(), t["has_obs"].data_ptr(),
                                       t["desc"].data_ptr())
    d_assign = torch.empty((n, cap), dtype=torch.int32, device=dev)
    d_res = torch.empty((3, n), dtype=torch.int32, device=dev)
    prm = views.make_track_params(W, H, th=1.0, nnratio=0.8)
    for stereo in (True, False):
        matcher.TrackLocalMapBatch_device(exl, n, oL["kps"].data_ptr(), oL["desc"].data_ptr(), oL["n"].data_ptr(), cap,
                                          d_ur.data_ptr() if stereo else 0, 0, d_fr.data_ptr(), dmap, 0, prm,
                                          d_assign.data_ptr(), d_res[0].data_ptr(), d_res[1].data_ptr(),
                                          d_res[2].data_ptr(), s)
        torch.cuda.synchronize()
        assign, res = d_assign.cpu().numpy(), d_res.cpu().numpy()
        assert not res[2].any()
        for i in range(n):
            r = dict(ref[i])
            nl = len(r["kl"])
            if not stereo:
                r["ur"] = None
            nm, a_ref, nv = _oracle_track(r, frs[i], maps[i], np.zeros(nl, np.uint8), 1.0, 0.8, False, 0.0)
            assert res[1][i] == nv and res[0][i] == nm and np.array_equal(assign[i, :nl], a_ref), (stereo, i)
    # a candidate list that is too small is reported per frame, never silently truncated
    small = views.make_track_params(W, H, th=15.0, nnratio=0.8, cand_per_frame=64)
    matcher.TrackLocalMapBatch_device(exl, n, oL["kps"].data_ptr(), oL["desc"].data_ptr(), oL["n"].data_ptr(), cap,
                                      d_ur.data_ptr(), 0, d_fr.data_ptr(), dmap, 0, small, d_assign.data_ptr(),
                                      d_res[0].data_ptr(), d_res[1].data_ptr(), d_res[2].data_ptr(), s)
    torch.cuda.synchronize()
    assert (d_res[2].cpu().numpy() == -2).all()
    with pytest.raises(OrbxError):
        matcher.StereoTrackFramesBatch(ORBextractor(NF, max_batch=2), ORBextractor(NF, max_batch=2), L[:2], R[:2], MBF, MB,
                                       frs[:2], views.make_local_map(**{k: v.cpu().numpy() for k, v in t.items()}), small)


def test_multi_gpu_c_entry_points_shard_and_equal_the_single_device_call(matcher):
    """orbx_extract_batch_multi / orbm_stereo_track_frames_batch_multi (the C-level sharder over the GPUs of one box, one
    host thread per device): with every visible device (and, on a one-GPU box, two handle sets on device 0, which
    exercises the same threads and offsets) the outputs must equal the single-device call word for word — for an uneven
    split, with the default map rule and with an explicit map_index."""
    import torch
    ndev = torch.cuda.device_count()
    devs = list(range(ndev)) if ndev > 1 else [0, 0]
    n = 7  # 7 pairs over 2+ handle sets: uneven blocks
    L, R, ref = _frames(n, 80)
    m = 1500
    frs = np.stack([synth.frustum(W, H, seed=500 + i) for i in range(n)])
    n_maps = 3
    maps = [synth.local_map_world(ref[i]["kl"], ref[i]["dl"], m, frs[i], seed=600 + i) for i in range(n_maps)]
    lm = views.make_local_map(**{k: np.stack([mp[k] for mp in maps]) for k in maps[0]})
    prm = views.make_track_params(W, H, th=3.0, nnratio=0.8)
    occ = (np.random.default_rng(2).random((n, NF + 16 * 8)) < 0.2).astype(np.uint8)
    exl1, exr1 = ORBextractor(NF, max_batch=2), ORBextractor(NF, max_batch=2)
    mts = [ORBmatcher(0.8, True, device=d) for d in devs]
    exls = [ORBextractor(NF, max_batch=2, device=d) for d in devs]
    exrs = [ORBextractor(NF, max_batch=2, device=d) for d in devs]
    for mi in (None, np.array([2, 2, 0, 1, 1, 0, 2], np.int32)):
        one = matcher.StereoTrackFramesBatch(exl1, exr1, L, R, MBF, MB, frs, lm, prm, map_index=mi, occupied=occ)
        many = ORBmatcher.StereoTrackFramesBatchMulti(mts, exls, exrs, L, R, MBF, MB, frs, lm, prm, map_index=mi,
                                                      occupied=occ)
        for k in ("n_l", "n_r", "n_matched", "nmatches", "n_in_view"):
            assert np.array_equal(one[k], many[k]), k
        for i in range(n):
            nl, nr = int(one["n_l"][i]), int(one["n_r"][i])
            for k in ("kps_l", "desc_l", "u_right", "depth", "assign"):
                assert one[k][i, :nl].tobytes() == many[k][i, :nl].tobytes(), (k, i)
            assert one["desc_r"][i, :nr].tobytes() == many["desc_r"][i, :nr].tobytes()
        assert int(one["nmatches"].sum()) > 40  # maps are anchored on frames 0..2 only: few but real matches
    stereo = ORBmatcher.StereoTrackFramesBatchMulti(mts, exls, exrs, L, R, MBF, MB)
    assert np.array_equal(stereo["n_matched"], one["n_matched"])
    n1, mono1, k1, d1 = exl1.extract_batch(L, (0, 0))
    n2, mono2, k2, d2 = ORBextractor.extract_batch_multi(exls, L, (0, 0))
    assert np.array_equal(n1, n2) and np.array_equal(mono1, mono2)
    assert all(k1[i, :n1[i]].tobytes() == k2[i, :n1[i]].tobytes() and d1[i, :n1[i]].tobytes() == d2[i, :n1[i]].tobytes()
               for i in range(n))


def test_recorded_small_call_replays_with_the_callers_current_data():
    """A one-group host call that comes back with the same buffers, sizes and parameters is recorded as a CUDA graph on
    its second run and replayed from the third on (stereo_frames_body, orbm_api.cu). The replay must read what the
    caller's buffers hold NOW (new images, poses, occupancy in the same pinned memory), a changed parameter must leave
    the recorded path, and pageable buffers must keep working — every result against the oracle."""
    import torch
    mt = ORBmatcher(0.8, True)
    n, m = 2, 1500
    exl, exr = ORBextractor(NF, max_batch=2), ORBextractor(NF, max_batch=2)
    cap = exl.capacity

    def pinned(shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory().numpy()[:nbytes].view(dtype).reshape(shape)

    imgs_l, imgs_r = pinned((n, H, W), np.uint8), pinned((n, H, W), np.uint8)
    frs = pinned((n,), views.FRUSTUM_DTYPE)
    occ = pinned((n, cap), np.uint8)
    out = ORBmatcher.alloc_track_outputs(n, cap, empty=pinned)
    sets = []
    for seed0 in (1200, 1300):
        L, R, ref = _frames(n, seed0)
        f = np.stack([synth.frustum(W, H, seed=seed0 + 50 + i) for i in range(n)])
        maps = [synth.local_map_world(ref[i]["kl"], ref[i]["dl"], m, f[i], seed=seed0 + 70 + i) for i in range(n)]
        o = (np.random.default_rng(seed0).random((n, cap)) < 0.2).astype(np.uint8)
        sets.append((L, R, ref, f, maps, o))
    # the maps live in ONE set of pinned arrays whose contents are swapped with the scene
    map_arrays = {k: pinned(np.stack([mp[k] for mp in sets[0][4]]).shape, sets[0][4][0][k].dtype) for k in sets[0][4][0]}
    lm = views.make_local_map(**map_arrays)

    def load(k):
        L, R, ref, f, maps, o = sets[k]
        imgs_l[:], imgs_r[:], frs[:], occ[:] = L, R, f, o
        for key in map_arrays:
            map_arrays[key][:] = np.stack([mp[key] for mp in maps])

    def check(k, th):
        L, R, ref, f, maps, o = sets[k]
        for i in range(n):
            r = ref[i]
            nl = len(r["kl"])
            assert int(out["n_l"][i]) == nl and _same(out["kps_l"][i, :nl], r["kl"]) and _same(out["desc_l"][i, :nl], r["dl"])
            assert int(out["n_matched"][i]) == r["nm"] and _same(out["u_right"][i, :nl], r["ur"])
            nm, assign, nv = _oracle_track(r, f[i], maps[i], o[i, :nl], th, 0.8, False, 0.0)
            assert int(out["n_in_view"][i]) == nv and int(out["nmatches"][i]) == nm
            assert np.array_equal(out["assign"][i, :nl], assign), "pair %d assign differs" % i

    mt._L.orbm_debug_graph_launches.argtypes = [C.c_void_p]
    launches = lambda: int(mt._L.orbm_debug_graph_launches(mt._h))
    prm = views.make_track_params(W, H, th=1.0, nnratio=0.8)
    load(0)
    for rep in range(4):  # direct, (direct again if a buffer grew), recorded, replayed
        for v in out.values():
            v[...] = 0
        mt.StereoTrackFramesBatch(exl, exr, imgs_l, imgs_r, MBF, MB, frs, lm, prm, occupied=occ, out=out)
        check(0, 1.0)
    assert launches() >= 1, "the repeated call never ran as a recorded graph"
    before = launches()
    load(1)  # same buffers, new contents
    mt.StereoTrackFramesBatch(exl, exr, imgs_l, imgs_r, MBF, MB, frs, lm, prm, occupied=occ, out=out)
    check(1, 1.0)
    assert launches() == before + 1
    prm3 = views.make_track_params(W, H, th=3.0, nnratio=0.8)  # another key: back on the direct path
    mt.StereoTrackFramesBatch(exl, exr, imgs_l, imgs_r, MBF, MB, frs, lm, prm3, occupied=occ, out=out)
    check(1, 3.0)
    assert launches() == before + 1
    # pageable buffers, repeated: whichever path the library takes, the words must be the oracle's
    L, R, ref, f, maps, o = sets[0]
    lm2 = views.make_local_map(**{k: np.stack([mp[k] for mp in maps]) for k in maps[0]})
    out2 = ORBmatcher.alloc_track_outputs(n, cap)
    Lc, Rc, fc, oc = L.copy(), R.copy(), f.copy(), o.copy()
    for rep in range(4):
        mt.StereoTrackFramesBatch(exl, exr, Lc, Rc, MBF, MB, fc, lm2, prm, occupied=oc, out=out2)
        out, keep = out2, out
        check(0, 1.0)
        out = keep
