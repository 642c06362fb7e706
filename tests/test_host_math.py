"""The product's __host__ __device__ arithmetic headers (csrc/orbx_math.h, orbx_plan.h, orbx_quadtree.h), compiled
for the CPU by tests/Makefile -> libhostcheck.so, against glibc / libstdc++ / cv2 / the oracle. These are the exact
sources the kernels compile, so a green run here means a GPU mismatch can only come from the kernels' data movement.
No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from orb_slam3_fast_b200 import synth
from oracle import orbref

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-C", HERE, "-s"])
    L = C.CDLL(os.path.join(HERE, "libhostcheck.so"))
    L.hc_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.hc_fast_atan2.restype = C.c_float
    L.hc_sincosf.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
    L.hc_cv_round.argtypes = [C.c_float]
    L.hc_std_sort_perm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.hc_std_sort_perm_warp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.hc_sincosf_sweep.argtypes = [C.c_uint32, C.c_uint32]
    L.hc_sincosf_sweep.restype = C.c_long
    L.hc_heap_calls.restype = C.c_long
    L.hc_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
    L.hc_axis_table.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hc_quadtree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_fast_atan2_matches_oracle_and_cv2(hc):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    pts = rng.integers(-300000, 300000, (30000, 2)).astype(np.float32)
    pts[:200] = rng.integers(-3, 4, (200, 2))
    for y, x in pts:
        a = np.float32(hc.hc_fast_atan2(y, x))
        assert a.tobytes() == np.float32(cv2.fastAtan2(float(y), float(x))).tobytes(), (y, x)


def test_sincosf_is_glibc_bit_exact(hc):
    # every float in the range the descriptor stage uses: angle * pi/180 in [0, 2*pi]  (src/ORBextractor.cc:106-107)
    lo = np.float32(0.0).view(np.uint32)
    hi = np.float32(6.3).view(np.uint32)
    # stride through the range in 64 slabs of 2^18 consecutive bit patterns (~1.7e7 values) + the low end exhaustively
    step = (int(hi) - int(lo)) // 64
    bad = 0
    for k in range(64):
        s = int(lo) + k * step
        bad += hc.hc_sincosf_sweep(s, s + (1 << 18))
    assert bad == 0
    # the 360 * 64 angles fastAtan2 can actually produce are dense near multiples of its polynomial grid: spot-check
    c, s = C.c_float(), C.c_float()
    libm = C.CDLL("libm.so.6")
    libm.cosf.argtypes = libm.sinf.argtypes = [C.c_float]
    libm.cosf.restype = libm.sinf.restype = C.c_float
    for deg in np.linspace(0, 360, 5000, dtype=np.float32):
        rad = np.float32(deg) * np.float32(np.pi / np.float32(180.0))
        hc.hc_sincosf(rad, C.byref(c), C.byref(s))
        assert np.float32(c.value).tobytes() == np.float32(libm.cosf(float(rad))).tobytes()
        assert np.float32(s.value).tobytes() == np.float32(libm.sinf(float(rad))).tobytes()


def test_logf_is_glibc_bit_exact(hc):
    """MapPoint::PredictScale's log(ratio) (src/MapPoint.cc:566) and mfLogScaleFactor (src/Frame.cc:189) are glibc logf; the
    device restatement must equal the host's for every input (1 in 64 of all non-negative float patterns here, plus the
    ranges PredictScale lives in exhaustively: ratio in [2^-8, 2^8])."""
    hc.hc_logf_sweep.restype = C.c_long
    hc.hc_logf_sweep.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    assert hc.hc_logf_sweep(0, 0x7f800001, 64) == 0
    lo = int(np.float32(2.0 ** -8).view(np.uint32))
    hi = int(np.float32(2.0 ** 8).view(np.uint32))
    assert hc.hc_logf_sweep(lo, hi, 1) == 0
    assert hc.hc_logf_sweep(0x80000000, 0x80000010, 1) == 0 and hc.hc_logf_sweep(0x7f800000, 0x7f800010, 1) == 0


def test_is_in_frustum_arithmetic_matches_oracle(hc):
    """The product's Frame::isInFrustum arithmetic (orbx_math.h, what k_frustum runs) against the oracle's restatement on
    the CPU: a GPU mismatch can then only come from data movement."""
    from orb_slam3_fast_b200 import synth
    img = synth.scene(480, 640, seed=11)
    _, kps, desc = orbref.Extractor(1200)(img, (0, 0))
    for seed in range(4):
        fr = synth.frustum(640, 480, seed=seed)
        mp = synth.local_map_world(kps, desc, 20000, fr, seed=seed)
        nv, o = orbref.is_in_frustum(fr, orbref.make_local_map(**mp))
        assert nv > 5000
        m = 20000
        got = dict(track_in_view=np.zeros(m, np.uint8), proj_x=np.zeros(m, np.float32), proj_y=np.zeros(m, np.float32),
                   proj_xr=np.zeros(m, np.float32), level=np.zeros(m, np.int32), view_cos=np.zeros(m, np.float32),
                   depth=np.zeros(m, np.float32))
        frb = np.ascontiguousarray(fr).reshape(1).view(np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        hc.hc_is_in_frustum.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_float] + [C.c_void_p] * 7
        hc.hc_is_in_frustum.restype = None
        arr = {k: np.ascontiguousarray(v) for k, v in mp.items()}
        hc.hc_is_in_frustum(p(frb), p(arr["pos"]), p(arr["normal"]), p(arr["min_dist"]), p(arr["max_dist"]),
                            p(arr["skip"]), m, 0.5, p(got["track_in_view"]), p(got["proj_x"]), p(got["proj_y"]),
                            p(got["proj_xr"]), p(got["level"]), p(got["view_cos"]), p(got["depth"]))
        for k in o:
            assert got[k].tobytes() == o[k].tobytes(), k


def test_cv_round_ties_to_even(hc):
    for v, r in ((0.5, 0), (1.5, 2), (2.5, 2), (-0.5, 0), (-1.5, -2), (-2.5, -2), (3.49, 3), (3.51, 4)):
        assert hc.hc_cv_round(v) == r


@pytest.mark.parametrize("which", ["serial", "warp"])
def test_std_sort_emulation_reproduces_libstdcxx_permutation(hc, which):
    sort = hc.hc_std_sort_perm if which == "serial" else hc.hc_std_sort_perm_warp
    # DistributeOctTree sorts (size, UL.x) pairs with std::sort (src/ORBextractor.cc:686); ties are common and the
    # permutation libstdc++'s introsort leaves them in decides which node is split last (SURVEY.md §7 hard part 1)
    rng = np.random.default_rng(5)
    for n in list(range(1, 40)) + [63, 64, 65, 100, 200, 333, 500, 1000]:
        for rep in range(6):
            k0 = rng.integers(2, 2 + max(2, n // 8), n).astype(np.int32)
            k1 = (rng.integers(0, 16, n) * 37).astype(np.int32)
            got = np.empty(n, np.int32)
            sort(_p(k0), _p(k1), n, _p(got))
            ref = orbref.std_sort_perm(k0, k1)  # the real std::sort, compiled into the oracle
            assert np.array_equal(got, ref), (n, rep)
    # fully random keys, all-equal keys, two-valued keys, sorted and reversed inputs
    for n in (17, 33, 64, 129, 300, 777, 2048):
        for k0 in (rng.integers(0, 1 << 18, n), np.zeros(n), rng.integers(0, 2, n), np.arange(n), np.arange(n)[::-1],
                   np.repeat(np.arange((n + 4) // 5), 5)[:n]):
            k0 = np.ascontiguousarray(k0, np.int32)
            k1 = np.zeros(n, np.int32)
            got = np.empty(n, np.int32)
            sort(_p(k0), _p(k1), n, _p(got))
            assert np.array_equal(got, orbref.std_sort_perm(k0, k1)), n
    # adversarial inputs that push introsort into its heapsort fallback must take that path too
    n = 2000
    k0 = np.arange(n, dtype=np.int32)[::-1].copy()
    k1 = np.zeros(n, np.int32)
    got = np.empty(n, np.int32)
    sort(_p(k0), _p(k1), n, _p(got))
    assert np.array_equal(got, orbref.std_sort_perm(k0, k1))


class LevelPlan(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("pitch", C.c_int), ("img_off", C.c_int64), ("maxBX", C.c_int),
                ("maxBY", C.c_int), ("nCols", C.c_int), ("nRows", C.c_int), ("wCell", C.c_int), ("hCell", C.c_int),
                ("cell_base", C.c_int), ("slot_cap", C.c_int), ("slot_base", C.c_int), ("quota", C.c_int),
                ("nIni", C.c_int), ("hX", C.c_float), ("kp_cap", C.c_int), ("kp_base", C.c_int), ("scale", C.c_float),
                ("inv_scale", C.c_float), ("sigma2", C.c_float), ("inv_sigma2", C.c_float), ("patch", C.c_int),
                ("xtab_off", C.c_int), ("ytab_off", C.c_int)]


class Plan(C.Structure):
    _fields_ = [("nlevels", C.c_int), ("w", C.c_int), ("h", C.c_int), ("cells_per_frame", C.c_int),
                ("slots_per_frame", C.c_int), ("kps_per_frame", C.c_int), ("pyr_bytes_per_frame", C.c_int64),
                ("tab_entries", C.c_int), ("max_tile_bytes", C.c_int), ("max_cell_px", C.c_int),
                ("max_quota", C.c_int), ("umax", C.c_int * 16), ("lv", LevelPlan * 16)]  # kMaxLevels


@pytest.mark.parametrize("w,h,nf,sf,nl", [(640, 480, 1000, 1.2, 8), (752, 480, 1200, 1.2, 8), (1280, 720, 2000, 1.2, 8),
                                          (640, 480, 1500, 1.1, 8), (640, 480, 300, 1.5, 3), (241, 241, 500, 1.2, 8)])
def test_plan_matches_oracle_setup(hc, w, h, nf, sf, nl):
    assert hc.hc_plan_size() == C.sizeof(Plan)
    P = Plan()
    assert hc.hc_plan(w, h, nf, sf, nl, C.byref(P)) == 0
    ex = orbref.Extractor(nf, sf, nl)
    ex(synth.uniform_noise(h, w, 0), (0, 0))
    assert list(P.umax) == ex.umax.tolist()
    for l in range(nl):
        L = P.lv[l]
        assert (L.w, L.h) == ex.level_dims(l)
        assert L.quota == ex.features_per_level[l]
        assert np.float32(L.scale) == ex.scale[l] and np.float32(L.inv_scale) == ex.inv_scale[l]
        assert np.float32(L.sigma2) == ex.sigma2[l] and np.float32(L.inv_sigma2) == ex.inv_sigma2[l]
        assert L.pitch >= L.w and L.pitch % 4 == 0


def test_plan_rejects_degenerate_sizes(hc):
    P = Plan()
    assert hc.hc_plan(100, 100, 1000, 1.2, 8, C.byref(P)) == -2   # a level without a 35-px cell (:781-784)
    assert hc.hc_plan(5000, 480, 1000, 1.2, 8, C.byref(P)) == -3
    assert hc.hc_plan(0, 480, 1000, 1.2, 8, C.byref(P)) == -1


def test_axis_table_drives_an_exact_resize(hc):
    cv2 = pytest.importorskip("cv2")
    img = synth.scene(200, 300, 2)
    dw, dh = 250, 167
    def tab(ss, ds, clamp):
        o, a, b = (np.empty(ds, np.int16) for _ in range(3))
        hc.hc_axis_table(ss, ds, clamp, _p(o), _p(a), _p(b))
        return o.astype(np.int64), a.astype(np.int64), b.astype(np.int64)
    xo, xa, xb = tab(300, dw, 1)
    yo, ya, yb = tab(200, dh, 0)
    src = img.astype(np.int64)
    x1 = np.minimum(xo + 1, 299)
    hrow = src[:, xo] * xa + src[:, x1] * xb                       # horizontal pass, every source row
    y0, y1 = np.clip(yo, 0, 199), np.clip(yo + 1, 0, 199)
    v = (((ya[:, None] * (hrow[y0] >> 4)) >> 16) + ((yb[:, None] * (hrow[y1] >> 4)) >> 16) + 2) >> 2
    got = np.clip(v, 0, 255).astype(np.uint8)
    assert np.array_equal(got, cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("kind,h,w,nf,seed", [("scene", 480, 640, 1000, 0), ("scene", 480, 752, 1200, 1),
                                              ("noise_blur", 480, 640, 1000, 2), ("uniform_noise", 480, 640, 1000, 3),
                                              ("scene", 300, 900, 300, 4), ("scene", 720, 1280, 2000, 5),
                                              ("scene", 480, 640, 5000, 6), ("scene", 241, 241, 50, 7)])
def test_array_quadtree_reproduces_distribute_octtree(hc, kind, h, w, nf, seed):
    # the warp-oriented array quadtree (csrc/orbx_quadtree.h) vs the oracle's std::list restatement of
    # DistributeOctTree (src/ORBextractor.cc:557-757), level by level on the oracle's own candidates
    img = synth.make(kind, h, w, seed)
    ex = orbref.Extractor(nf)
    ex(img, (0, 0))
    P = Plan()
    assert hc.hc_plan(w, h, nf, 1.2, 8, C.byref(P)) == 0
    for l in range(8):
        L = P.lv[l]
        cand = ex.level_candidates(l)
        want = ex.level_keypoints(l)
        packed = (cand["x"].astype(np.uint32) | (cand["y"].astype(np.uint32) << 12) |
                  (cand["response"].astype(np.uint32) << 24)).astype(np.uint32)
        packed = np.ascontiguousarray(packed)
        out = np.zeros(L.kp_cap + 8, np.uint32)
        n = hc.hc_quadtree(_p(packed) if len(packed) else None, len(packed), L.maxBX - 16, L.maxBY - 16, L.nIni, L.hX,
                           L.quota, L.kp_cap, _p(out))
        assert n == len(want), "level %d: %d vs %d" % (l, n, len(want))
        sel = cand[out[:n]]
        assert np.array_equal(sel["x"] + 16, want["x"]) and np.array_equal(sel["y"] + 16, want["y"]), "level %d" % l
        assert np.array_equal(sel["response"], want["response"])
