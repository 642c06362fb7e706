"""First-light diagnostic for the GPU box: runs one frame, reports per-stage agreement with the oracle into
gpurun_out/first_light.txt (so that one gpurun call tells which stage to look at)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam3_fast_b200 import ORBextractor, synth  # noqa: E402
from oracle import orbref  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "first_light.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    out.write(s + "\n")
    out.flush()


def run(kind, h, w, nf, lap):
    img = synth.make(kind, h, w, 3)
    ex, ex_ref = ORBextractor(nf), orbref.Extractor(nf)
    t = time.time()
    mono, kps, desc = ex(img, lap)
    P("== %s %dx%d nf=%d lap=%s: n=%d mono=%d (%.1f ms incl. setup)" % (kind, w, h, nf, lap, len(kps), mono,
                                                                       (time.time() - t) * 1e3))
    mono_r, kps_r, desc_r = ex_ref(img, lap)
    P("   oracle: n=%d mono=%d" % (len(kps_r), mono_r))
    for l in range(8):
        raw, raw_r = ex.debug_level(l), ex_ref.level_image(l)
        c, c_r = ex.debug_candidates(l), ex_ref.level_candidates(l)
        k, k_r = ex.debug_level_keypoints(l), ex_ref.level_keypoints(l)
        b, b_r = ex.debug_level(l, True), ex_ref.level_blurred(l)
        same_c = len(c) == len(c_r) and all(np.array_equal(c[f], c_r[f]) for f in ("x", "y", "response"))
        same_k = len(k) == len(k_r) and all(np.array_equal(k[f], k_r[f]) for f in ("x", "y", "response"))
        P("   L%d raw diff px=%d | cand %d vs %d same=%s | kp %d vs %d same=%s | blur diff px=%s" %
          (l, (raw != raw_r).sum(), len(c), len(c_r), same_c, len(k), len(k_r), same_k,
           "n/a" if b_r is None else (b != b_r).sum()))
        if not same_c and len(c) and len(c_r):
            n = min(len(c), len(c_r))
            d = np.nonzero((c["x"][:n] != c_r["x"][:n]) | (c["y"][:n] != c_r["y"][:n]) |
                           (c["response"][:n] != c_r["response"][:n]))[0]
            P("      first cand diffs at", d[:5], "gpu", c[d[:3]], "ref", c_r[d[:3]])
    if len(kps) == len(kps_r):
        for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
            P("   final %s mismatches: %d" % (f, (kps[f] != kps_r[f]).sum()))
        P("   final desc rows differing: %d" % (desc != desc_r).any(axis=1).sum())
        if (kps["angle"] != kps_r["angle"]).any():
            i = np.nonzero(kps["angle"] != kps_r["angle"])[0][:5]
            P("   angle gpu", kps["angle"][i], "ref", kps_r["angle"][i])


try:
    run("scene", 480, 640, 1000, (0, 0))
    run("scene", 480, 752, 1200, (0, 1000))
    run("uniform_noise", 480, 640, 1000, (0, 0))
except Exception:
    P(traceback.format_exc())
