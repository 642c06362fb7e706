"""Writes tools/ubench/data/track_case.bin: bench.py's configs[3] inputs (D distinct 640x480 stereo pairs, one pose, one
10 000-point local map and one occupancy array per pair) together with the CPU oracle's outputs for them, so that
tools/ubench/track_check (no Python, seconds of GPU box time) can gate parity and time the batched tracking path on
the box. Run here, before gpurun; the file is git-ignored but travels with the snapshot.

  layout (little endian): int32 magic 'ORBT', D, w, h, nfeatures, M, cap; float32 mbf, mb; orbx_track_params (40 B);
      u8 left[D][h][w], right[D][h][w]; orbx_frustum[D] (104 B each);
      pos[D][M][3] f32, normal[D][M][3] f32, min_dist[D][M] f32, max_dist[D][M] f32, skip[D][M] u8, has_obs[D][M] u8,
      desc[D][M][32] u8; occupied[D][cap] u8;
      per pair: for eye in (L, R): int32 n, mono; kps[n] (28 B each); desc[n][32]
                then int32 n_matched; f32 u_right[nL]; f32 depth[nL]; int32 nmatches, n_in_view; int32 assign[nL]
"""
import ctypes as C
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from orb_slam3_fast_b200 import views  # noqa: E402
from oracle import orbref  # noqa: E402


def main():
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    cfg = bench.CONFIGS[3]
    w, h, nf, M = cfg["w"], cfg["h"], cfg["nfeat"], cfg["map_points"]
    cap = nf + 16 * bench.NLEVELS
    L, R = bench.make_pairs(D, 300, w, h)
    rl = orbref.Extractor(nf, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH)
    rr = orbref.Extractor(nf, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH)
    ext = []
    for i in range(D):
        ml, kl, dl = rl(L[i], (0, 0))
        mr, kr, dr = rr(R[i], (0, 0))
        nm, ur, dp = orbref.stereo_match(rl, rr, kl, dl, kr, dr, bench.MBF, bench.MB)
        ext.append((ml, kl, dl, mr, kr, dr, nm, ur, dp))
    frs, stacked = bench.make_track_inputs(cfg, L, 300, [(e[1], e[2]) for e in ext])
    occ = (np.random.default_rng(5).random((D, cap)) < 0.25).astype(np.uint8)
    prm = views.make_track_params(w, h, th=cfg["th"], nnratio=cfg["nnratio"])
    out = os.path.join(ROOT, "tools", "ubench", "data", "track_case.bin")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "wb") as f:
        f.write(struct.pack("<7i2f", 0x4F524254, D, w, h, nf, M, cap, bench.MBF, bench.MB))
        f.write(bytes(prm))
        f.write(L.tobytes())
        f.write(R.tobytes())
        f.write(np.ascontiguousarray(frs).tobytes())
        for k in ("pos", "normal", "min_dist", "max_dist", "skip", "has_obs", "desc"):
            f.write(np.ascontiguousarray(stacked[k]).tobytes())
        f.write(occ.tobytes())
        for i in range(D):
            ml, kl, dl, mr, kr, dr, nm, ur, dp = ext[i]
            for mono, k, d in ((ml, kl, dl), (mr, kr, dr)):
                f.write(struct.pack("<2i", len(k), mono))
                f.write(np.ascontiguousarray(k).tobytes())
                f.write(np.ascontiguousarray(d).tobytes())
            f.write(struct.pack("<i", nm))
            f.write(ur.astype(np.float32).tobytes())
            f.write(dp.astype(np.float32).tobytes())
            nl = len(kl)
            off, items = orbref.build_grid(kl, 0.0, 0.0, prm.inv_w, prm.inv_h)
            g, keep = orbref.make_grid(off, items, 0.0, 0.0, prm.inv_w, prm.inv_h)
            fv = orbref.make_frame_view(kl, dl, ur, occ[i][:nl], g, keep, rl.scale)
            lm1 = orbref.make_local_map(**{k2: v[i] for k2, v in stacked.items()})
            tn, ta, tv = orbref.track_local_map(fv, frs[i], lm1, 0, cfg["th"], cfg["nnratio"])
            f.write(struct.pack("<2i", tn, tv))
            f.write(np.ascontiguousarray(ta, np.int32).tobytes())
    print("%s: %d pairs, %.1f MB" % (out, D, os.path.getsize(out) / 1e6))


if __name__ == "__main__":
    main()
