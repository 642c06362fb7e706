"""Development aid: per-phase cycle accounting of k_quadtree (needs liborbx.so built with -DORBX_QT_PROF)."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam3_fast_b200 import ORBextractor, synth
ex = ORBextractor(1200, max_batch=8)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 752
imgs = np.stack([synth.stereo_pair(480, W, s)[0] for s in range(8)])
ex.extract_batch(imgs)
ex.extract_batch(imgs)
names = ["count", "gather", "init", "select", "sort", "walk", "children", "prep", "sweep", "tail", "rounds", "rounds2",
         "C", "n"]
ex._L.orbx_debug_qt_profile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
for f in (0, 3):
    for l in range(8):
        buf = np.zeros(16, np.int64)
        rc = ex._L.orbx_debug_qt_profile(ex._h, f, l, buf.ctypes.data_as(C.c_void_p))
        if rc:
            print("rc", rc); break
        tot = buf[:10].sum()
        print("f%d l%d total %6.1f us | " % (f, l, tot / 1965.0) +
              " ".join("%s=%s" % (n, ("%.1f" % (v / 1965.0)) if i < 10 else int(v)) for i, (n, v) in enumerate(zip(names, buf))))
