#!/usr/bin/env python
"""Freezes tests/golden/*.npz — TEST INFRASTRUCTURE.

The reference holds no tests / golden vectors for this path and cannot be built here, so the
goldens are frozen from the strongest independent source this image has: the REAL OpenCV 4.13 kernels driven through
cv2 by oracle/cv2_pipeline.py (orchestration restated from src/ORBextractor.cc serial path; resize / FAST / blur /
fastAtan2 are OpenCV's own, cosf/sinf glibc's own, std::sort libstdc++'s own) and cv2.BFMatcher for knn2.

Inputs are regenerated from the seeded generators in orb_slam3_fast_b200/synth.py; every file also stores the CRC32 of
its input so a drifting generator is detected instead of silently comparing different images.

Run from the repo root in the build container:  python tools/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from orb_slam3_fast_b200 import synth  # noqa: E402
from oracle import cv2_pipeline  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name: (kind, h, w, seed, nfeatures, scale, nlevels, iniTh, minTh, lapping)
EXTRACT_CASES = {
    "cfg1_640x480_mono": ("scene", 480, 640, 0, 1000, 1.2, 8, 20, 7, (0, 1000)),      # BASELINE configs[0]
    "cfg1_640x480_lap00": ("scene", 480, 640, 0, 1000, 1.2, 8, 20, 7, (0, 0)),
    "cfg2_752x480_left": ("stereo_l", 480, 752, 0, 1200, 1.2, 8, 20, 7, (0, 0)),       # BASELINE configs[1]
    "cfg2_752x480_right": ("stereo_r", 480, 752, 0, 1200, 1.2, 8, 20, 7, (0, 0)),
    "cfg3_1280x720_mono": ("scene", 720, 1280, 2, 2000, 1.2, 8, 20, 7, (0, 1000)),     # BASELINE configs[2]
    "noise_blur_640x480": ("noise_blur", 480, 640, 3, 1000, 1.2, 8, 20, 7, (0, 0)),
    "low_contrast12_640x480": ("low_contrast12", 480, 640, 1, 1000, 1.2, 8, 20, 7, (0, 0)),  # no corner at all: N = 0
    "low_contrast40_640x480": ("low_contrast40", 480, 640, 1, 1000, 1.2, 8, 20, 7, (0, 0)),  # minThFAST retry in most cells
    "min_size_241": ("scene", 241, 241, 3, 500, 1.2, 8, 20, 7, (0, 0)),
    "wide_900x300_lap": ("scene", 300, 900, 3, 300, 1.2, 8, 20, 7, (100, 500)),        # 3 quadtree roots
    "params_1500_1p1": ("scene", 480, 640, 11, 1500, 1.1, 8, 15, 5, (0, 0)),
}


def make_image(kind, h, w, seed):
    if kind == "stereo_l":
        return synth.stereo_pair(h, w, seed)[0]
    if kind == "stereo_r":
        return synth.stereo_pair(h, w, seed)[1]
    if kind.startswith("low_contrast"):
        return synth.low_contrast(h, w, seed, amplitude=int(kind[len("low_contrast"):]))
    return synth.make(kind, h, w, seed)


def main():
    import cv2
    os.makedirs(GOLD, exist_ok=True)
    for name, (kind, h, w, seed, nf, sf, nl, it, mt, lap) in EXTRACT_CASES.items():
        img = make_image(kind, h, w, seed)
        ex = cv2_pipeline.Cv2Extractor(nf, sf, nl, it, mt)
        mono, kps, desc = ex(img, lap)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), crc=np.uint32(zlib.crc32(img.tobytes())),
                            mono=np.int32(mono), kps=kps, desc=desc,
                            level_crc=np.array([zlib.crc32(p.tobytes()) for p in ex.pyramid], np.uint32))
        print("%-24s n=%d mono=%d" % (name, len(kps), mono))
    # knn2: cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) (the call at src/Frame.cc:1293), incl. a tie-heavy case
    for name, nq, nt, proto in (("knn2_1k", 1000, 1000, 0), ("knn2_ties", 700, 900, 32)):
        q, t = synth.descriptors(nq, 5, proto), synth.descriptors(nt, 6, proto)
        mm = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), crc=np.uint32(zlib.crc32(q.tobytes() + t.tobytes())),
                            idx1=np.array([m[0].trainIdx for m in mm], np.int32),
                            d1=np.array([int(m[0].distance) for m in mm], np.int32),
                            idx2=np.array([m[1].trainIdx for m in mm], np.int32),
                            d2=np.array([int(m[1].distance) for m in mm], np.int32))
        print(name)


if __name__ == "__main__":
    main()
