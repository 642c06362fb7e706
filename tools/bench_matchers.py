#!/usr/bin/env python
"""Per-call timings of the matcher rows of SURVEY.md §8 (a11-a16) on 1 x B200, each next to the CPU oracle, plus
BASELINE.json configs[4] (the brute-force 256-bit Hamming 2-NN sweep 1k x 1k ... 100k x 100k).
Parity with the oracle / numpy is asserted before anything is timed. Usage: python tools/bench_matchers.py > out.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, synth, views  # noqa: E402
from oracle import orbref  # noqa: E402  (checker + CPU baseline only)


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return 1e3 * (time.perf_counter() - t0) / reps


def knn2_sweep(mt):
    rows = []
    for n in (1000, 3000, 10000, 30000, 100000):
        q, t = synth.descriptors(n, 21), synth.descriptors(n, 22)
        dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
        i1 = torch.empty(n, dtype=torch.int32, device="cuda"); d1 = torch.empty_like(i1)
        i2 = torch.empty_like(i1); d2 = torch.empty_like(i1)
        st = torch.cuda.current_stream().cuda_stream

        def dev():
            mt.knnMatch2_device(dq.data_ptr(), n, dt.data_ptr(), n, i1.data_ptr(), d1.data_ptr(), i2.data_ptr(),
                                d2.data_ptr(), st)
        for _ in range(2):
            dev()
        reps = 20 if n <= 10000 else 3
        torch.cuda.synchronize()   # the call runs on the matcher's own stream: device-wide syncs bracket the region
        t0 = time.perf_counter()
        for _ in range(reps):
            dev()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / reps
        # spot check against numpy on 16 rows
        rr = np.random.default_rng(n).integers(0, n, 16)
        D = np.bitwise_count(q[rr].view(np.uint64)[:, None, :] ^ t.view(np.uint64)[None, :, :]).sum(axis=2)
        order = np.lexsort((np.broadcast_to(np.arange(n), D.shape), D), axis=1)[:, :2]
        assert np.array_equal(i1.cpu().numpy()[rr], order[:, 0]) and np.array_equal(i2.cpu().numpy()[rr], order[:, 1])
        row = {"n": n, "ms_device": ms, "gpairs_per_s": n * n / ms / 1e6,
               "ms_host_call": timed(lambda: mt.knnMatch2(q, t), 3 if n > 10000 else 10)}
        # A/B: the same call with the tensor-core path switched off (every size on the POPC kernel)
        os.environ["ORBM_KNN2_TC"] = "0"
        dev()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dev()
        torch.cuda.synchronize()
        del os.environ["ORBM_KNN2_TC"]
        row["ms_device_popc_kernel"] = 1e3 * (time.perf_counter() - t0) / reps
        if n <= 3000:
            import cv2
            bf = cv2.BFMatcher(cv2.NORM_HAMMING)
            row["ms_cv2_bfmatcher_host_threads"] = timed(lambda: bf.knnMatch(q, t, k=2), 3)
        rows.append(row)
    return rows


def main():
    mt = ORBmatcher(0.8)
    out = {"knn2_sweep_configs4": knn2_sweep(mt)}
    # ---- a12: Frame::ComputeStereoMatches on one 752x480 pair (per-call, host keypoints in / uRight, depth out) ----
    w, h = 752, 480
    left, right, _ = synth.stereo_pair(h, w, 5)
    exl, exr = ORBextractor(1200), ORBextractor(1200)
    _, kl, dl = exl(left)
    _, kr, dr = exr(right)
    mbf, mb = float(np.float32(435.2 * 0.11)), float(np.float32(0.11))
    rl, rr_ = orbref.Extractor(1200), orbref.Extractor(1200)
    rl(left, (0, 0)); rr_(right, (0, 0))
    nm, u, d = mt.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, mbf, mb)
    _, u_r, d_r = orbref.stereo_match(rl, rr_, kl, dl, kr, dr, mbf, mb)
    assert np.array_equal(u, u_r) and np.array_equal(d, d_r)
    out["stereo_match_one_pair"] = {
        "ms_gpu_call": timed(lambda: mt.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, mbf, mb), 50),
        "ms_cpu_oracle": timed(lambda: orbref.stereo_match(rl, rr_, kl, dl, kr, dr, mbf, mb), 10), "matched": int(nm)}
    # ---- a14: SearchByProjection(Frame&, const Frame&) after the caller-side projection ----
    ur = np.where(np.random.default_rng(0).random(len(kl)) < 0.7, kl["x"] - 10, -1).astype(np.float32)
    occ = np.zeros(len(kl), np.uint8)
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    off, items = views.assign_features_to_grid(kl, 0.0, 0.0, inv_w, inv_h)
    fv = views.make_frame_view(kl, dl, ur, occ, off, items, 0.0, 0.0, inv_w, inv_h, exl.GetScaleFactors())
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    fr = orbref.make_frame_view(kl, dl, ur, occ, g, keep, exl.GetScaleFactors())
    pts = synth.projected_points(kl, dl, 1200, w, h, 8, exl.GetScaleFactors(), 0, th=7.0, stereo=True)
    m2 = ORBmatcher(0.9, True)
    pv, pr = views.make_projected(**pts), orbref.make_projected(**pts)
    n, a = m2.SearchByProjectionProjected(fv, pv, 100)
    n_r, a_r = orbref.search_by_projection_frame(fr, pr, 100, True)
    assert n == n_r and np.array_equal(a, a_r)
    out["search_by_projection_frame_1200pts"] = {
        "ms_gpu_call": timed(lambda: m2.SearchByProjectionProjected(fv, pv, 100), 50),
        "ms_cpu_oracle": timed(lambda: orbref.search_by_projection_frame(fr, pr, 100, True), 20), "matches": int(n)}
    # ---- §8f rank 3: MapPoint::ComputeDistinctiveDescriptors for 10 000 map points (2..30 observations each) ----
    rng = np.random.default_rng(1)
    lists = []
    for _ in range(10000):
        n = int(rng.integers(2, 31))
        lists.append(synth.descriptors(n, int(rng.integers(1 << 30))))
    best = mt.ComputeDistinctiveDescriptors(lists)
    assert all(int(best[k]) == orbref.distinctive_descriptor(lists[k]) for k in range(0, 10000, 37))
    out["distinctive_descriptors_10k_points"] = {
        "ms_gpu_call": timed(lambda: mt.ComputeDistinctiveDescriptors(lists), 5),
        "ms_cpu_oracle": timed(lambda: [orbref.distinctive_descriptor(d) for d in lists], 2),
        "observations": int(sum(len(d) for d in lists)),
        "note": "the GPU figure includes the Python-side concatenation of the 10 000 lists and the H2D / D2H copies"}
    # ---- §8f rank 2 / 3 rows on one 752x480 stereo pair's features (1500 per image) ----
    e1, e2 = ORBextractor(1500), ORBextractor(1500)
    _, k1, d1 = e1(left)
    _, k2, d2 = e2(right)
    rngf = np.random.default_rng(4)

    def featvec(desc):
        node_of = (desc[:, 0].astype(np.int64) >> 2) * 7 + 3      # 64 "vocabulary nodes"
        ids, inv = np.unique(node_of, return_inverse=True)
        order = np.argsort(inv, kind="stable")
        offsets = np.zeros(len(ids) + 1, np.int32)
        offsets[1:] = np.cumsum(np.bincount(inv, minlength=len(ids)))
        return ids.astype(np.uint32), offsets, order.astype(np.uint32)
    sf, s2 = e1.GetScaleFactors(), e1.GetScaleSigmaSquares()
    vg, vr = [], []
    for k, d in ((k1, d1), (k2, d2)):
        ids, off, idx = featvec(d)
        a = (k, d, np.full(len(k), -1, np.float32), (rngf.random(len(k)) < 0.6).astype(np.uint8), ids, off, idx, sf, s2)
        vg.append(views.make_keyframe_view(*a))
        vr.append(orbref.make_keyframe_view(*a))
    mb_ = ORBmatcher(0.7, True)
    n, mf = mb_.SearchByBoW(vg[0], vg[1])
    n_r, mf_r = orbref.search_by_bow(vr[0], vr[1], 0.7, True)
    assert n == n_r and np.array_equal(mf, mf_r)
    out["search_by_bow_kf_frame"] = {"ms_gpu_call": timed(lambda: mb_.SearchByBoW(vg[0], vg[1]), 30),
                                     "ms_cpu_oracle": timed(lambda: orbref.search_by_bow(vr[0], vr[1], 0.7, True), 10),
                                     "matches": int(n)}
    voc = synth.vocabulary(10, 5, 0, ragged=False)              # 111 111 nodes
    mt.SetVocabulary(views.make_vocabulary(**voc))
    vref = orbref.make_vocabulary(**voc)
    w, wt, nd = mt.BowTransform(d1, 4)
    w_r, wt_r, nd_r = orbref.bow_transform(vref, d1, 4)
    assert np.array_equal(w, w_r) and np.array_equal(nd, nd_r)
    out["bow_transform_1500_features_depth5"] = {"ms_gpu_call": timed(lambda: mt.BowTransform(d1, 4), 30),
                                                 "ms_cpu_oracle": timed(lambda: orbref.bow_transform(vref, d1, 4), 10),
                                                 "vocabulary_nodes": int(len(voc["descriptors"]))}
    inv_w2, inv_h2 = np.float32(64) / np.float32(w_img := 752), np.float32(48) / np.float32(480)
    off2, items2 = views.assign_features_to_grid(k1, 0.0, 0.0, inv_w2, inv_h2)
    kfv = views.make_frame_view(k1, d1, None, np.zeros(len(k1), np.uint8), off2, items2, 0.0, 0.0, inv_w2, inv_h2, sf)
    g2, keep2 = orbref.make_grid(off2, items2, 0.0, 0.0, inv_w2, inv_h2)
    kfr = orbref.make_frame_view(k1, d1, None, np.zeros(len(k1), np.uint8), g2, keep2, sf)
    mpts = 5000
    srcp = rngf.integers(0, len(k1), mpts)
    lev = np.clip(k1["octave"][srcp] + rngf.integers(-1, 2, mpts), 0, 7).astype(np.int32)
    fp = dict(u=(k1["x"][srcp] + rngf.normal(0, 1.2, mpts)).astype(np.float32),
              v=(k1["y"][srcp] + rngf.normal(0, 1.2, mpts)).astype(np.float32), u_right=None,
              radius=(np.float32(3.0) * np.asarray(sf, np.float32)[lev]).astype(np.float32), min_level=lev - 1,
              max_level=lev, angle=np.zeros(mpts, np.float32), has_obs=np.zeros(mpts, np.uint8),
              desc=synth.flip_bits(d1[srcp], rngf.integers(0, 60, mpts), rngf))
    inv_s2 = 1.0 / np.asarray(s2, np.float32)
    pg, pr2 = views.make_projected(**fp), orbref.make_projected(**fp)
    bi, bd = mt.FuseMatch(kfv, inv_s2, pg)
    bi_r, bd_r = orbref.fuse_match(kfr, inv_s2, pr2)
    assert np.array_equal(bi, bi_r) and np.array_equal(bd, bd_r)
    out["fuse_match_5000_points"] = {"ms_gpu_call": timed(lambda: mt.FuseMatch(kfv, inv_s2, pg), 30),
                                     "ms_cpu_oracle": timed(lambda: orbref.fuse_match(kfr, inv_s2, pr2), 10),
                                     "fused": int((bd_r <= 50).sum())}
    out["timer"] = "host wall clock around synchronous ABI calls unless the key says device (CUDA events)"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
