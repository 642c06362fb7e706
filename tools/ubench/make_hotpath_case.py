"""Writes tools/ubench/data/hotpath_case.bin: the bench's synthetic stereo pairs (bench.make_pairs(n, 0)) together with
the CPU oracle's outputs for them (keypoints, descriptors, mvuRight, mvDepth), so that tools/ubench/hotpath_check (no
Python, seconds of GPU box time) can gate parity and time the stages on the box. Run here, before gpurun; the file is
git-ignored but travels with the snapshot.

  layout (little endian): int32 magic, n_pairs, w, h, nfeatures; float32 mbf, mb;
                          u8 left[n_pairs][h][w], right[n_pairs][h][w];
                          per pair: for eye in (L, R): int32 n, mono; kps[n] (28 B each); desc[n][32]
                                    then int32 n_matched; float32 u_right[nL]; float32 depth[nL]
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import orbref  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    L, R = bench.make_pairs(n_pairs, 0)
    out = os.path.join(ROOT, "tools", "ubench", "data", "hotpath_case.bin")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    rl = orbref.Extractor(bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH)
    rr = orbref.Extractor(bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH)
    with open(out, "wb") as f:
        f.write(struct.pack("<5i2f", 0x4F524258, n_pairs, bench.W, bench.H, bench.NFEAT, bench.MBF, bench.MB))
        f.write(L.tobytes())
        f.write(R.tobytes())
        for i in range(n_pairs):
            ml, kl, dl = rl(L[i], (0, 0))
            mr, kr, dr = rr(R[i], (0, 0))
            for mono, k, d in ((ml, kl, dl), (mr, kr, dr)):
                f.write(struct.pack("<2i", len(k), mono))
                f.write(np.ascontiguousarray(k).tobytes())
                f.write(np.ascontiguousarray(d).tobytes())
            nm, ur, dp = orbref.stereo_match(rl, rr, kl, dl, kr, dr, bench.MBF, bench.MB)
            f.write(struct.pack("<i", nm))
            f.write(ur.astype(np.float32).tobytes())
            f.write(dp.astype(np.float32).tobytes())
    print("%s: %d pairs, %.1f MB" % (out, n_pairs, os.path.getsize(out) / 1e6))


if __name__ == "__main__":
    main()
