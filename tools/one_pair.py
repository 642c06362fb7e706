"""Development aid: ONE 640x480 stereo pair + a 10 000-point local map through orbm_stereo_track_frames_batch, a few times
(for `ncu --metrics gpu__time_duration.sum`: the per-kernel times of the single-pair call)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, synth, views
from oracle import orbref
W, H, NF = 640, 480, 1200
left, right, _ = synth.stereo_pair(H, W, 3)
L, R = left[None], right[None]
exl, exr = ORBextractor(NF, max_batch=1), ORBextractor(NF, max_batch=1)
_, kl, dl = orbref.Extractor(NF)(left, (0, 0))
fr = np.stack([synth.frustum(W, H, seed=7)])
mp = synth.local_map_world(kl, dl, 10000, fr[0], seed=8)
lm = views.make_local_map(**{k: v[None] for k, v in mp.items()})
prm = views.make_track_params(W, H, th=1.0, nnratio=0.8)
mt = ORBmatcher(0.8, True)
mbf, mb = float(np.float32(435.2 * 0.11)), float(np.float32(0.11))
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = mt.StereoTrackFramesBatch(exl, exr, L, R, mbf, mb, fr, lm, prm)
print("ok", int(out["n_l"][0]), int(out["nmatches"][0]))
