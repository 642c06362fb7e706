"""ctypes loader of oracle/_ref/liborbref_src.so: the reference's OWN src/ORBextractor.cc, compiled where it lies under
/root/reference against the stand-in OpenCV / TBB headers of oracle/ref_stubs (image primitives = the oracle's
cv2-pinned ones, TBB = serial). TEST INFRASTRUCTURE: it checks the oracle's restatement against the reference code."""
import ctypes as C
import os

import numpy as np

from oracle.orbref import KP_DTYPE

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "liborbref_src.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.orbrefsrc_create.restype = C.c_void_p
        L.orbrefsrc_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbrefsrc_destroy.argtypes = [C.c_void_p]
        L.orbrefsrc_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orbrefsrc_extract.restype = C.c_int
        L.orbrefsrc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orbrefsrc_distribute.restype = C.c_int
        L.orbrefsrc_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
        _lib = L
    return _lib


class ReferenceExtractor:
    """ORB_SLAM3::ORBextractor of the reference source."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = lib().orbrefsrc_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orbrefsrc_destroy(self._h)
            self._h = None

    def tables(self):
        f = [np.empty(self.nlevels, np.float32) for _ in range(4)]
        per_level, umax = np.empty(self.nlevels, np.int32), np.empty(16, np.int32)
        lib().orbrefsrc_tables(self._h, *[a.ctypes.data for a in f], per_level.ctypes.data, umax.ctypes.data)
        return f[0], f[1], f[2], f[3], per_level, umax

    def __call__(self, img, lapping=(0, 0)):
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps, desc = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        if img is None or img.size == 0:
            mono = lib().orbrefsrc_extract(self._h, None, 0, 0, 0, lapping[0], lapping[1], kps.ctypes.data,
                                           desc.ctypes.data, cap, C.byref(n))
        else:
            img = np.ascontiguousarray(img, np.uint8)
            mono = lib().orbrefsrc_extract(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                           lapping[0], lapping[1], kps.ctypes.data, desc.ctypes.data, cap, C.byref(n))
        if mono == -1000:
            raise RuntimeError("capacity")
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def keypoints(self, img, serial_twin):
        """Per-level keypoints after ComputePyramid + ComputeKeyPointsOctTree (serial twin) or its TBB twin."""
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps, counts = np.zeros(cap, KP_DTYPE), np.zeros(self.nlevels, np.int32)
        fn = lib().orbrefsrc_keypoints
        fn.restype = C.c_int
        n = fn(C.c_void_p(self._h), C.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0],
               int(serial_twin), C.c_void_p(kps.ctypes.data), cap, C.c_void_p(counts.ctypes.data))
        if n < 0:
            raise RuntimeError("capacity")
        return kps[:n].copy(), counts

    def distribute(self, kps, min_x, max_x, min_y, max_y, n_want, level):
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        out = np.zeros(len(kps) + 8, KP_DTYPE)
        n = lib().orbrefsrc_distribute(self._h, kps.ctypes.data, len(kps), min_x, max_x, min_y, max_y, n_want, level,
                                       out.ctypes.data, len(out))
        if n < 0:
            raise RuntimeError("capacity")
        return out[:n].copy()


# ---- the reference's own src/ORBmatcher.cc (oracle/_ref/liborbref_matcher_src.so) ---------------------------------
_MPATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "liborbref_matcher_src.so")
_mlib = None


def matcher_available():
    return os.path.exists(_MPATH)


def mlib():
    global _mlib
    if _mlib is None:
        _mlib = C.CDLL(_MPATH)
        for name in ("orbrefsrc_descriptor_distance", "orbrefsrc_search_by_projection_map",
                     "orbrefsrc_search_for_triangulation", "orbrefsrc_search_by_bow", "orbrefsrc_search_by_bow_kf",
                     "orbrefsrc_search_by_bow_fisheye", "orbrefsrc_stereo_fisheye", "orbrefsrc_search_by_bow_kf_fisheye", "orbrefsrc_search_for_triangulation_fisheye",
                     "orbrefsrc_search_for_initialization", "orbrefsrc_search_by_projection_last_frame",
                     "orbrefsrc_search_by_projection_keyframe", "orbrefsrc_fuse", "orbrefsrc_fuse_two_camera",
                     "orbrefsrc_search_by_projection_last_frame_fisheye",
                     "orbrefsrc_search_by_projection_sim3", "orbrefsrc_search_by_sim3", "orbrefsrc_stereo_frame",
                     "orbrefsrc_features_in_area", "orbrefsrc_distinctive_descriptor"):
            getattr(_mlib, name).restype = C.c_int
    return _mlib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return mlib().orbrefsrc_descriptor_distance(_p(a), _p(b))


def search_by_projection_map(fv, mps, th, nnratio, far_points=False, th_far=0.0):
    assign = np.empty(max(fv.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_projection_map(fv.ref(), mps.ref(), C.c_float(th), C.c_float(nnratio),
                                                  int(far_points), C.c_float(th_far), _p(assign))
    return n, assign[:fv.struct.n]


def is_in_frustum(frustum, local_map, map_index=0, viewing_cos_limit=0.5, out=None):
    """The reference's own Frame::isInFrustum + MapPoint::PredictScale text on the stand-in world."""
    from . import orbref
    fn = mlib().orbrefsrc_is_in_frustum
    fn.restype = None
    return orbref.is_in_frustum(frustum, local_map, map_index, viewing_cos_limit, out, fn=fn)


def search_local_points(fv, frustum, local_map, held, bad, ctl, th_far, state, map_index=0):
    """Tracking::SearchLocalPoints (src/Tracking.cc:3249-3330) — the reference's own function text on a stand-in Tracking
    object (in a shim world: the drop-in body of shim/Tracking_orbx.cc). held[n]: what mvpMapPoints[i] holds on entry
    (-1 nothing, k >= 0 local-map point k, -2 a point outside the local map); bad[m]; ctl = (mSensor, isImuInitialized,
    GetIniertialBA2, mState, frame id, mnLastRelocFrameId, mbFarPoints); state: dict of the per-point words on entry
    (track_in_view, proj_x, proj_y, proj_xr, level, view_cos, depth, visible, last_seen). Returns (assign[n], state on
    return, project_points[m][2] with NaN where mmProjectPoints has no entry, len(mmProjectPoints))."""
    n, m = fv.struct.n, local_map.struct.m
    fr = np.ascontiguousarray(frustum).reshape(1)
    held = np.ascontiguousarray(held, np.int32)
    bad = np.ascontiguousarray(bad, np.uint8)
    c = np.zeros(8, np.int32)
    c[:len(ctl)] = ctl
    st = {k: np.ascontiguousarray(v).copy() for k, v in state.items()}
    assign = np.full(max(n, 1), -7, np.int32)
    pp = np.zeros((max(m, 1), 2), np.float32)
    fn = mlib().orbrefsrc_search_local_points
    fn.restype = C.c_int
    k = fn(fv.ref(), _p(fr), local_map.ref(), C.c_int(map_index), _p(held), _p(bad), _p(c), C.c_float(th_far), _p(assign),
           _p(st["track_in_view"]), _p(st["proj_x"]), _p(st["proj_y"]), _p(st["proj_xr"]), _p(st["level"]),
           _p(st["view_cos"]), _p(st["depth"]), _p(st["visible"]), _p(st["last_seen"]), _p(pp))
    return assign[:n], st, pp[:m], k


def search_by_projection_map_fisheye(fv, mps, mr, th, nnratio, far_points=False, th_far=0.0):
    """The reference's own SearchByProjection(Frame&, vector<MapPoint*>) on a two-camera stand-in Frame."""
    from . import orbref
    return orbref.search_by_projection_map_fisheye(fv, mps, mr, th, nnratio, far_points, th_far,
                                                   fn=mlib().orbrefsrc_search_by_projection_map_fisheye)


def search_for_triangulation(kf1, kf2, F12, ep, only_stereo=False, coarse=False, check_orientation=True):
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    m = np.empty(max(kf1.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_for_triangulation(kf1.ref(), kf2.ref(), _p(F), C.c_float(ep[0]), C.c_float(ep[1]),
                                                  int(only_stereo), int(coarse), int(check_orientation), _p(m))
    return n, m[:kf1.struct.n]


def search_for_triangulation_fisheye(kf1, n_left1, kf2, n_left2, F12x4, only_stereo=False, coarse=False,
                                     check_orientation=True):
    """The reference's own SearchForTriangulation on two-camera stand-in KeyFrames."""
    from . import orbref
    mlib().orbrefsrc_search_for_triangulation_fisheye.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                                  C.c_int, C.c_int, C.c_int, C.c_void_p]
    return orbref.search_for_triangulation_fisheye(kf1, n_left1, kf2, n_left2, F12x4, only_stereo, coarse,
                                                   check_orientation, fn=mlib().orbrefsrc_search_for_triangulation_fisheye)


def search_by_bow(kf, frame, nnratio=0.7, check_orientation=True):
    m = np.empty(max(frame.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_bow(kf.ref(), frame.ref(), C.c_float(nnratio), int(check_orientation), _p(m))
    return n, m[:frame.struct.n]


def search_by_bow_fisheye(kf, n_left_kf, frame, n_left_f, nnratio=0.7, check_orientation=True):
    """The reference's own SearchByBoW(KeyFrame*, Frame&, ...) with pKF->NLeft = n_left_kf, F.Nleft = n_left_f and both
    second cameras set (the keypoints of rows >= NLeft live in mvKeysRight)."""
    m = np.empty(max(frame.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_bow_fisheye(kf.ref(), int(n_left_kf), frame.ref(), int(n_left_f), C.c_float(nnratio),
                                               int(check_orientation), _p(m))
    return n, m[:frame.struct.n]


def search_by_bow_kf(kf1, kf2, nnratio=0.8, check_orientation=True):
    m = np.empty(max(kf1.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_bow_kf(kf1.ref(), kf2.ref(), C.c_float(nnratio), int(check_orientation), _p(m))
    return n, m[:kf1.struct.n]


def stereo_fisheye(kps_l, desc_l, kps_r, desc_r, mono_left, mono_right, level_sigma2, Rlr, tlr):
    """Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1271-1331) with the stand-in KannalaBrandt8: returns
    (n, mvLeftToRightMatch, mvRightToLeftMatch, mvDepth, mvStereo3Dpoints)."""
    nl, nr = len(kps_l), len(kps_r)
    kl, kr = np.ascontiguousarray(kps_l), np.ascontiguousarray(kps_r)
    dl, dr = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
    s2 = np.ascontiguousarray(level_sigma2, np.float32)
    R, t = np.ascontiguousarray(Rlr, np.float32).reshape(9), np.ascontiguousarray(tlr, np.float32).reshape(3)
    l2r, r2l = np.empty(max(nl, 1), np.int32), np.empty(max(nr, 1), np.int32)
    depth, p3d = np.empty(max(nl, 1), np.float32), np.zeros((max(nl, 1), 3), np.float32)
    n = mlib().orbrefsrc_stereo_fisheye(_p(kl), _p(dl), nl, _p(kr), _p(dr), nr, int(mono_left), int(mono_right), _p(s2),
                                        len(s2), _p(R), _p(t), _p(l2r), _p(r2l), _p(depth), _p(p3d))
    return n, l2r[:nl], r2l[:nr], depth[:nl], p3d[:nl]


def search_by_bow_kf_fisheye(kf1, n_un1, kf2, n_un2, nnratio=0.8, check_orientation=True):
    """The reference's own SearchByBoW(KeyFrame*, KeyFrame*, ...) with NLeft != -1 and mvKeysUn of n_un rows."""
    m = np.empty(max(kf1.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_bow_kf_fisheye(kf1.ref(), int(n_un1), kf2.ref(), int(n_un2), C.c_float(nnratio),
                                                  int(check_orientation), _p(m))
    return n, m[:kf1.struct.n]


def search_for_initialization(f1, f2, prev_xy, window_size=100, nnratio=0.9, check_orientation=True):
    prev = np.ascontiguousarray(prev_xy, np.float32).reshape(-1, 2)
    m = np.empty(max(f1.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_for_initialization(f1.ref(), f2.ref(), _p(prev), int(window_size), C.c_float(nnratio),
                                                   int(check_orientation), _p(m))
    return n, m[:f1.struct.n]


def search_by_projection_last_frame(fv, u, v, z, octave, angle, has_obs, desc, th, mbf, mb, mode, check_orientation=True):
    a = [np.ascontiguousarray(x, t) for x, t in ((u, np.float32), (v, np.float32), (z, np.float32), (octave, np.int32),
                                                 (angle, np.float32), (has_obs, np.uint8), (desc, np.uint8))]
    assign = np.empty(max(fv.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_projection_last_frame(fv.ref(), len(a[0]), *[_p(x) for x in a], C.c_float(th),
                                                         C.c_float(mbf), C.c_float(mb), int(mode),
                                                         int(check_orientation), _p(assign))
    return n, assign[:fv.struct.n]


def search_by_projection_last_frame_fisheye(fisheye_view, u, v, z, octave, angle, has_obs, desc, th, mbf, mb, mode,
                                            check_orientation=True):
    """SearchByProjection(CurrentFrame, LastFrame, th, false) on a two-camera current frame; assign[n_left + n_right]."""
    a = [np.ascontiguousarray(x, t) for x, t in ((u, np.float32), (v, np.float32), (z, np.float32), (octave, np.int32),
                                                 (angle, np.float32), (has_obs, np.uint8), (desc, np.uint8))]
    n_rows = fisheye_view.struct.n_left + fisheye_view.struct.n_right
    assign = np.empty(max(n_rows, 1), np.int32)
    n = mlib().orbrefsrc_search_by_projection_last_frame_fisheye(fisheye_view.ref(), len(a[0]), *[_p(x) for x in a],
                                                                 C.c_float(th), C.c_float(mbf), C.c_float(mb), int(mode),
                                                                 int(check_orientation), _p(assign))
    return n, assign[:n_rows]


def search_by_projection_keyframe(fv, u, v, level, angle, found, desc, th, orb_dist, check_orientation=True):
    a = [np.ascontiguousarray(x, t) for x, t in ((u, np.float32), (v, np.float32), (level, np.int32),
                                                 (angle, np.float32), (found, np.uint8), (desc, np.uint8))]
    assign = np.empty(max(fv.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_projection_keyframe(fv.ref(), len(a[0]), *[_p(x) for x in a], C.c_float(th),
                                                       int(orb_dist), int(check_orientation), _p(assign))
    return n, assign[:fv.struct.n]


def fuse(kfv, inv_level_sigma2, u, v, z, level, desc, th, mbf, sim3=False):
    a = [np.ascontiguousarray(x, t) for x, t in ((u, np.float32), (v, np.float32), (z, np.float32), (level, np.int32),
                                                 (desc, np.uint8))]
    s2 = np.ascontiguousarray(inv_level_sigma2, np.float32)
    best = np.empty(max(len(a[0]), 1), np.int32)
    n = mlib().orbrefsrc_fuse(kfv.ref(), _p(s2), len(a[0]), *[_p(x) for x in a], C.c_float(th), C.c_float(mbf),
                              int(sim3), _p(best))
    return n, best[:len(a[0])]


def fuse_two_camera(fisheye_view, inv_level_sigma2, u, v, z, level, desc, th, mbf, b_right):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight) on a two-camera KeyFrame; best[i] = row of mDescriptors or -1."""
    a = [np.ascontiguousarray(x, t) for x, t in ((u, np.float32), (v, np.float32), (z, np.float32), (level, np.int32),
                                                 (desc, np.uint8))]
    s2 = np.ascontiguousarray(inv_level_sigma2, np.float32)
    best = np.empty(max(len(a[0]), 1), np.int32)
    n = mlib().orbrefsrc_fuse_two_camera(fisheye_view.ref(), _p(s2), len(a[0]), *[_p(x) for x in a], C.c_float(th),
                                         C.c_float(mbf), int(bool(b_right)), _p(best))
    return n, best[:len(a[0])]


def search_by_projection_sim3(kfv, matched_in, u, v, level, desc, th, ratio_hamming, with_kfs=False):
    a = [np.ascontiguousarray(x, t) for x, t in ((u, np.float32), (v, np.float32), (level, np.int32), (desc, np.uint8))]
    mi = np.ascontiguousarray(matched_in, np.uint8)
    assign = np.empty(max(kfv.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_projection_sim3(kfv.ref(), _p(mi), len(a[0]), *[_p(x) for x in a], int(th),
                                                   C.c_float(ratio_hamming), int(with_kfs), _p(assign))
    return n, assign[:kfv.struct.n]


def search_by_sim3(v1, v2, side1, side2, th):
    """side = (has_mappoint u8[n], u f32[n], v f32[n], level i32[n], desc u8[n, 32]) per KeyFrame."""
    args = []
    for has, u, v, level, desc in (side1, side2):
        args += [np.ascontiguousarray(has, np.uint8), np.ascontiguousarray(u, np.float32),
                 np.ascontiguousarray(v, np.float32), np.ascontiguousarray(level, np.int32),
                 np.ascontiguousarray(desc, np.uint8)]
    m = np.empty(max(v1.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_search_by_sim3(v1.ref(), v2.ref(), *[_p(x) for x in args], C.c_float(th), _p(m))
    return n, m[:v1.struct.n]


# ---- DBoW2's own TemplatedVocabulary<FORB> as vendored by the reference (oracle/_ref/liborbref_dbow2_src.so) --------
_VPATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "liborbref_dbow2_src.so")
_vlib = None


def vocabulary_available():
    return os.path.exists(_VPATH)


def vlib():
    global _vlib
    if _vlib is None:
        _vlib = _vocabulary_lib(_VPATH)
    return _vlib


def write_vocabulary_text(voc, k, path):
    """The flat tree of synth.vocabulary in the text format TemplatedVocabulary::loadFromTextFile reads: header
    'k L scoring weighting' (L1_NORM, TF_IDF), then one line per node in id order: parent, is_leaf, 32 bytes, weight."""
    off, children = voc["child_offsets"], voc["children"]
    n = len(off) - 1
    parent = np.zeros(n, np.int64)
    for p in range(n):
        parent[children[off[p]:off[p + 1]]] = p
    lines = ["%d %d 0 0" % (k, voc["depth"])]
    for i in range(1, n):
        leaf = int(off[i + 1] == off[i])
        lines.append("%d %d %s %s" % (parent[i], leaf, " ".join(str(int(b)) for b in voc["descriptors"][i]),
                                      repr(float(voc["weight"][i]))))
    with open(path, "w") as f:
        f.write("\n".join(lines))          # no trailing newline: the loader would read one more (empty) node


def _vocabulary_lib(path):
    lib = C.CDLL(path)
    lib.orbrefsrc_voc_load.restype = C.c_void_p
    lib.orbrefsrc_voc_load.argtypes = [C.c_char_p]
    lib.orbrefsrc_voc_destroy.argtypes = [C.c_void_p]
    lib.orbrefsrc_voc_size.argtypes = [C.c_void_p]
    lib.orbrefsrc_voc_transform.restype = C.c_int
    lib.orbrefsrc_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
    lib.orbrefsrc_compute_bow.restype = C.c_int
    lib.orbrefsrc_compute_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int]
    return lib


class ReferenceVocabulary:
    """DBoW2's own vocabulary object. `world`: path of another library of the same world (a shim bow world, where
    Frame / KeyFrame::ComputeBoW are the drop-in bodies of shim/FrameBoW_orbx.cc); the object is created, used and
    destroyed by that one library."""

    def __init__(self, path, world=None):
        self._lib = vlib() if world is None else _vocabulary_lib(world)
        self._h = self._lib.orbrefsrc_voc_load(path.encode())
        if not self._h:
            raise RuntimeError("loadFromTextFile failed")

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orbrefsrc_voc_destroy(self._h)
            self._h = None

    def words(self):
        return self._lib.orbrefsrc_voc_size(self._h)

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word, weight, node = np.empty(n, np.uint32), np.empty(n, np.float64), np.empty(n, np.uint32)
        bw, bvals = np.empty(n + 1, np.uint32), np.empty(n + 1, np.float64)
        k = self._lib.orbrefsrc_voc_transform(self._h, _p(desc), n, levelsup, _p(word), _p(weight), _p(node), _p(bw),
                                              _p(bvals), n + 1)
        return word, weight, node, bw[:k], bvals[:k]

    def compute_bow(self, desc, which=0):
        """Frame::ComputeBoW (which 0; 3: mBowVec already filled) / KeyFrame::ComputeBoW (1; 2: mBowVec filled, mFeatVec
        empty) on a stand-in object holding `desc`: (node_id[n] of mFeatVec, words, values of mBowVec)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        node = np.empty(max(n, 1), np.uint32)
        bw, bvals = np.empty(n + 2, np.uint32), np.empty(n + 2, np.float64)
        k = self._lib.orbrefsrc_compute_bow(self._h, _p(desc), n, which, _p(node), _p(bw), _p(bvals), n + 2)
        return node[:n], bw[:k], bvals[:k]


def stereo_frame(left, right, mbf, mb, nfeatures=1200, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """Both ORBextractor::operator() calls + Frame::ComputeStereoMatches, all of it the reference's own code.
    Returns (n_matched, kps_l, desc_l, kps_r, desc_r, u_right, depth)."""
    left, right = np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(right, np.uint8)
    cap = nfeatures + 8 * nlevels + 64
    kl, kr = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dl, dr = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    ur, dp = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    nl, nr = C.c_int(0), C.c_int(0)
    n = mlib().orbrefsrc_stereo_frame(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th, _p(left), _p(right),
                                      left.shape[1], left.shape[0], left.strides[0], C.c_float(mbf), C.c_float(mb),
                                      _p(kl), _p(dl), C.byref(nl), _p(kr), _p(dr), C.byref(nr), _p(ur), _p(dp), cap)
    if n == -1000:
        raise RuntimeError("capacity")
    return (n, kl[:nl.value].copy(), dl[:nl.value].copy(), kr[:nr.value].copy(), dr[:nr.value].copy(),
            ur[:nl.value].copy(), dp[:nl.value].copy())


class TrackParams(C.Structure):
    _fields_ = [("viewing_cos_limit", C.c_float), ("th", C.c_float), ("nnratio", C.c_float), ("far_points", C.c_int32),
                ("th_far", C.c_float), ("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float),
                ("inv_h", C.c_float), ("cand_per_frame", C.c_int32)]


def stereo_track_frame(left, right, mbf, mb, frustum, local_map, map_index, occupied, params, nfeatures=1200,
                       scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """configs[3] for one pair with the reference's own code: extract x2 + ComputeStereoMatches + AssignFeaturesToGrid +
    isInFrustum over the local map + SearchByProjection. `params`: any ctypes struct with the orbx_track_params layout.
    Returns dict(n_matched, kps_l, desc_l, kps_r, desc_r, u_right, depth, assign, nmatches, n_in_view)."""
    left, right = np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(right, np.uint8)
    cap = nfeatures + 8 * nlevels + 64
    kl, kr = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dl, dr = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    ur, dp = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    assign = np.full(cap, -1, np.int32)
    nl, nr, nm, nv = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    fr = np.ascontiguousarray(frustum).reshape(1)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    fn = mlib().orbrefsrc_stereo_track_frame
    fn.restype = C.c_int
    n = fn(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th, _p(left), _p(right), left.shape[1], left.shape[0],
           left.strides[0], C.c_float(mbf), C.c_float(mb), _p(fr), local_map.ref(), C.c_int(map_index),
           None if occ is None else _p(occ), C.byref(params), _p(kl), _p(dl), C.byref(nl), _p(kr), _p(dr), C.byref(nr),
           _p(ur), _p(dp), cap, _p(assign), C.byref(nm), C.byref(nv))
    if n == -1000:
        raise RuntimeError("capacity")
    return dict(n_matched=n, kps_l=kl[:nl.value], desc_l=dl[:nl.value], kps_r=kr[:nr.value], desc_r=dr[:nr.value],
                u_right=ur[:nl.value], depth=dp[:nl.value], assign=assign[:nl.value], nmatches=nm.value,
                n_in_view=nv.value)


def build_grid(fv):
    off, items = np.zeros(64 * 48 + 1, np.int32), np.zeros(max(fv.struct.n, 1), np.int32)
    mlib().orbrefsrc_build_grid(fv.ref(), _p(off), _p(items))
    return off, items[:off[-1]]


def features_in_area(fv, x, y, r, min_level=-1, max_level=-1, keyframe=False):
    out = np.empty(max(fv.struct.n, 1), np.int32)
    n = mlib().orbrefsrc_features_in_area(fv.ref(), C.c_float(x), C.c_float(y), C.c_float(r), int(min_level),
                                          int(max_level), int(keyframe), _p(out))
    return out[:n].copy()


def distinctive_descriptor(desc):
    """MapPoint::ComputeDistinctiveDescriptors over the rows of desc: the chosen descriptor (32 bytes) or None."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    out = np.zeros(32, np.uint8)
    ok = mlib().orbrefsrc_distinctive_descriptor(_p(desc), len(desc), _p(out))
    return out if ok else None
