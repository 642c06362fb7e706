#!/usr/bin/env python
"""BASELINE.json configs[3]: 640x480 stereo extract (1200 features) + SearchByProjection against a 10 000-MapPoint
synthetic local map, 1 x B200. Times, per frame: the extractor call, SearchByProjection through the host frame view
(orbm_search_by_projection_map), the device-resident form (orbm_search_by_projection_map_resident: grid built on the
device) and the CPU oracle. Parity of both GPU forms against the oracle is asserted first.
Usage (GPU box): python tools/bench_config4.py > gpurun_out/config4.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, synth, views  # noqa: E402
from oracle import orbref  # noqa: E402  (checker + CPU baseline only)


def main():
    w, h, m, th = 640, 480, 10000, 1.0
    ex = ORBextractor(1200)
    mt = ORBmatcher(0.8)
    img = synth.scene(h, w, 3)
    _, kps, desc = ex(img)
    rng = np.random.default_rng(3)
    ur = np.where(rng.random(len(kps)) < 0.7, kps["x"] - rng.uniform(1, 40, len(kps)), -1).astype(np.float32)
    occ = np.zeros(len(kps), np.uint8)
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    mp = synth.local_map(kps, desc, m, w, h, 8, 3)
    gp = (0.0, 0.0, inv_w, inv_h)

    def host_view():
        off, items = views.assign_features_to_grid(kps, 0.0, 0.0, inv_w, inv_h)
        fv = views.make_frame_view(kps, desc, ur, occ, off, items, 0.0, 0.0, inv_w, inv_h, ex.GetScaleFactors())
        return mt.SearchByProjection(fv, views.make_mappoints(**mp), th, True, 15.0)

    def resident():
        return mt.SearchByProjectionResident(ex, 0, len(kps), views.make_mappoints(**mp), gp, ur, occ, th, True, 15.0)

    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    fr = orbref.make_frame_view(kps, desc, ur, occ, g, keep, ex.GetScaleFactors())
    mpr = orbref.make_mappoints(**mp)

    def oracle():
        orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
        return orbref.search_by_projection_map(fr, mpr, th, 0.8, True, 15.0)

    n_o, a_o = oracle()
    for fn in (host_view, resident):
        n, a = fn()
        assert n == n_o and np.array_equal(a, a_o), fn.__name__

    def timed(fn, reps):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return 1e3 * (time.perf_counter() - t0) / reps

    out = {"workload": "configs[3]: 640x480, 1200 features, SearchByProjection vs 10000 MapPoints (th=1, nnratio 0.8)",
           "keypoints": int(len(kps)), "matches": int(n_o), "parity": "both GPU forms == oracle (assign[], count)",
           "ms_extract_host_call": timed(lambda: ex(img), 50),
           "ms_search_host_view": timed(host_view, 50),
           "ms_search_resident": timed(resident, 50),
           "ms_search_cpu_oracle_1thread": timed(oracle, 10),
           "timer": "host wall clock around the synchronous ABI calls (H2D of the 10k-point map and D2H of assign[] "
                    "included; the resident form skips the frame's H2D and the host grid)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
