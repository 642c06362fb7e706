"""ctypes binding of the CPU oracle (oracle/liborbref.so). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package (orb_slam3_fast_b200/) never does. See oracle/orbref.h for what each entry point restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborbref.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

GRID_COLS, GRID_ROWS = 64, 48


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("orbref.cpp", "orbref_mt.cpp", "orbref.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class Vocabulary(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("depth", C.c_int32), ("child_offsets", C.c_void_p), ("children", C.c_void_p),
                ("descriptors", C.c_void_p), ("word_id", C.c_void_p), ("weight", C.c_void_p)]


class Grid(C.Structure):
    _fields_ = [("cell_offsets", C.c_void_p), ("cell_items", C.c_void_p), ("min_x", C.c_float), ("min_y", C.c_float),
                ("inv_w", C.c_float), ("inv_h", C.c_float)]


class FrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("kps", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p),
                ("occupied", C.c_void_p), ("grid", Grid), ("scale_factors", C.c_void_p), ("n_levels", C.c_int32)]


class MapPoints(C.Structure):
    _fields_ = [("m", C.c_int32), ("track_in_view", C.c_void_p), ("proj_x", C.c_void_p), ("proj_y", C.c_void_p),
                ("proj_xr", C.c_void_p), ("level", C.c_void_p), ("view_cos", C.c_void_p), ("depth", C.c_void_p),
                ("has_obs", C.c_void_p), ("desc", C.c_void_p)]


class Projected(C.Structure):
    _fields_ = [("m", C.c_int32), ("u", C.c_void_p), ("v", C.c_void_p), ("u_right", C.c_void_p),
                ("radius", C.c_void_p), ("min_level", C.c_void_p), ("max_level", C.c_void_p), ("angle", C.c_void_p),
                ("has_obs", C.c_void_p), ("desc", C.c_void_p)]


class FeatVec(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_ids", C.c_void_p), ("offsets", C.c_void_p), ("indices", C.c_void_p)]


class KeyFrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("kps", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p),
                ("has_mappoint", C.c_void_p), ("featvec", FeatVec), ("scale_factors", C.c_void_p),
                ("level_sigma2", C.c_void_p), ("n_levels", C.c_int32)]


class Frustum(C.Structure):
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("fx", C.c_float),
                ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float), ("min_x", C.c_float),
                ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float), ("log_scale_factor", C.c_float),
                ("n_levels", C.c_int32)]


class LocalMap(C.Structure):
    _fields_ = [("m", C.c_int32), ("n_maps", C.c_int32), ("pos", C.c_void_p), ("normal", C.c_void_p),
                ("min_dist", C.c_void_p), ("max_dist", C.c_void_p), ("skip", C.c_void_p), ("has_obs", C.c_void_p),
                ("desc", C.c_void_p)]


class FisheyeView(C.Structure):
    _fields_ = [("n_left", C.c_int32), ("n_right", C.c_int32), ("kps_left", C.c_void_p), ("kps_right", C.c_void_p),
                ("desc", C.c_void_p), ("occupied", C.c_void_p), ("grid_left", Grid), ("grid_right", Grid),
                ("left_to_right", C.c_void_p), ("right_to_left", C.c_void_p), ("scale_factors", C.c_void_p),
                ("n_levels", C.c_int32)]


class MapPointsRight(C.Structure):
    _fields_ = [("track_in_view_r", C.c_void_p), ("proj_x_r", C.c_void_p), ("proj_y_r", C.c_void_p),
                ("level_r", C.c_void_p), ("view_cos_r", C.c_void_p)]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Holder:
    """Keeps numpy arrays alive next to the ctypes struct that points into them."""

    def __init__(self, struct, keep):
        self.struct = struct
        self.keep = keep

    def ref(self):
        return C.byref(self.struct)


def make_grid(offsets, items, min_x, min_y, inv_w, inv_h):
    offsets = _c(offsets, np.int32)
    items = _c(items, np.int32)
    g = Grid(_ptr(offsets), _ptr(items), float(min_x), float(min_y), float(inv_w), float(inv_h))
    return g, (offsets, items)


def make_frame_view(kps, desc, u_right, occupied, grid, grid_keep, scale_factors):
    kps = _c(kps, KP_DTYPE)
    desc = _c(desc, np.uint8)
    u_right = None if u_right is None else _c(u_right, np.float32)
    occupied = _c(occupied, np.uint8)
    sf = _c(scale_factors, np.float32)
    fv = FrameView(len(kps), _ptr(kps), _ptr(desc), _ptr(u_right), _ptr(occupied), grid, _ptr(sf), len(sf))
    return Holder(fv, (kps, desc, u_right, occupied, grid_keep, sf))


def make_mappoints(track_in_view, proj_x, proj_y, proj_xr, level, view_cos, depth, has_obs, desc):
    arrs = (_c(track_in_view, np.uint8), _c(proj_x, np.float32), _c(proj_y, np.float32), _c(proj_xr, np.float32),
            _c(level, np.int32), _c(view_cos, np.float32), _c(depth, np.float32), _c(has_obs, np.uint8),
            _c(desc, np.uint8))
    return Holder(MapPoints(len(arrs[0]), *[_ptr(a) for a in arrs]), arrs)


def make_local_map(pos, normal, min_dist, max_dist, skip, has_obs, desc):
    pos = _c(pos, np.float32)
    if pos.ndim == 2:
        pos = pos[None]
    n_maps, m = pos.shape[0], pos.shape[1]
    arrs = (pos, _c(normal, np.float32).reshape(n_maps, m, 3), _c(min_dist, np.float32).reshape(n_maps, m),
            _c(max_dist, np.float32).reshape(n_maps, m), None if skip is None else _c(skip, np.uint8).reshape(n_maps, m),
            _c(has_obs, np.uint8).reshape(n_maps, m), _c(desc, np.uint8).reshape(n_maps, m, 32))
    return Holder(LocalMap(m, n_maps, *[_ptr(a) for a in arrs]), arrs)


def _frustum_out(m, out):
    if out is None:
        out = dict(track_in_view=np.zeros(m, np.uint8), proj_x=np.zeros(m, np.float32), proj_y=np.zeros(m, np.float32),
                   proj_xr=np.zeros(m, np.float32), level=np.zeros(m, np.int32), view_cos=np.zeros(m, np.float32),
                   depth=np.zeros(m, np.float32))
    return out


def is_in_frustum(frustum, local_map, map_index=0, viewing_cos_limit=0.5, out=None, fn=None):
    """Frame::isInFrustum over local map `map_index` (orbref_is_in_frustum). frustum: a 104-byte record (numpy)."""
    fr = np.ascontiguousarray(frustum).reshape(1)
    assert fr.dtype.itemsize == 104
    out = _frustum_out(local_map.struct.m, out)
    fn = fn or lib().orbref_is_in_frustum
    fn(_ptr(fr), local_map.ref(), C.c_int(map_index), C.c_float(viewing_cos_limit), _ptr(out["track_in_view"]),
       _ptr(out["proj_x"]), _ptr(out["proj_y"]), _ptr(out["proj_xr"]), _ptr(out["level"]), _ptr(out["view_cos"]),
       _ptr(out["depth"]))
    return int(out["track_in_view"].sum()), out


def track_local_map(frame_view, frustum, local_map, map_index, th, nnratio, far_points=False, th_far=0.0,
                    viewing_cos_limit=0.5):
    """Tracking::SearchLocalPoints (src/Tracking.cc:3288-3330): isInFrustum over the local map, then SearchByProjection.
    Returns (nmatches, assign, n_in_view)."""
    m = local_map.struct.m
    nv, o = is_in_frustum(frustum, local_map, map_index, viewing_cos_limit)
    base = map_index * m
    has_obs = local_map.keep[5].reshape(-1)[base:base + m]
    desc = local_map.keep[6].reshape(-1, 32)[base:base + m]
    mps = make_mappoints(o["track_in_view"], o["proj_x"], o["proj_y"], o["proj_xr"], o["level"], o["view_cos"],
                         o["depth"], has_obs, desc)
    if nv == 0:
        return 0, np.full(frame_view.struct.n, -1, np.int32), 0
    nm, assign = search_by_projection_map(frame_view, mps, th, nnratio, far_points, th_far)
    return nm, assign, nv


def make_fisheye_view(kps_left, kps_right, desc, occupied, min_x, min_y, inv_w, inv_h, left_to_right, right_to_left,
                      scale_factors):
    """orbx_fisheye_view; the two grids are built here with orbref_build_grid (Frame::AssignFeaturesToGrid)."""
    kl, kr = _c(kps_left, KP_DTYPE), _c(kps_right, KP_DTYPE)
    desc, occupied = _c(desc, np.uint8), _c(occupied, np.uint8)
    l2r, r2l = _c(left_to_right, np.int32), _c(right_to_left, np.int32)
    sf = _c(scale_factors, np.float32)
    gl, keep_l = make_grid(*build_grid(kl, min_x, min_y, inv_w, inv_h), min_x, min_y, inv_w, inv_h)
    gr, keep_r = make_grid(*build_grid(kr, min_x, min_y, inv_w, inv_h), min_x, min_y, inv_w, inv_h)
    v = FisheyeView(len(kl), len(kr), _ptr(kl), _ptr(kr), _ptr(desc), _ptr(occupied), gl, gr, _ptr(l2r), _ptr(r2l),
                    _ptr(sf), len(sf))
    return Holder(v, (kl, kr, desc, occupied, keep_l, keep_r, l2r, r2l, sf))


def make_mappoints_right(track_in_view_r, proj_x_r, proj_y_r, level_r, view_cos_r):
    arrs = (_c(track_in_view_r, np.uint8), _c(proj_x_r, np.float32), _c(proj_y_r, np.float32), _c(level_r, np.int32),
            _c(view_cos_r, np.float32))
    return Holder(MapPointsRight(*[_ptr(a) for a in arrs]), arrs)


def search_by_projection_map_fisheye(fv, mps, mr, th, nnratio, far_points=False, th_far=0.0, fn=None):
    n_slots = fv.struct.n_left + fv.struct.n_right
    assign = np.empty(max(n_slots, 1), np.int32)
    fn = fn or lib().orbref_search_by_projection_map_fisheye
    fn.restype = C.c_int
    n = fn(fv.ref(), mps.ref(), mr.ref(), C.c_float(th), C.c_float(nnratio), C.c_int(int(far_points)), C.c_float(th_far),
           _ptr(assign))
    return n, assign[:n_slots]


def make_projected(u, v, u_right, radius, min_level, max_level, angle, has_obs, desc):
    arrs = (_c(u, np.float32), _c(v, np.float32), None if u_right is None else _c(u_right, np.float32),
            _c(radius, np.float32), _c(min_level, np.int32), _c(max_level, np.int32), _c(angle, np.float32),
            _c(has_obs, np.uint8), _c(desc, np.uint8))
    return Holder(Projected(len(arrs[0]), *[_ptr(a) for a in arrs]), arrs)


def make_keyframe_view(kps, desc, u_right, has_mappoint, node_ids, offsets, indices, scale_factors, level_sigma2):
    kps = _c(kps, KP_DTYPE)
    desc = _c(desc, np.uint8)
    u_right = None if u_right is None else _c(u_right, np.float32)
    hm = _c(has_mappoint, np.uint8)
    node_ids = _c(node_ids, np.uint32)
    offsets = _c(offsets, np.int32)
    indices = _c(indices, np.uint32)
    sf = _c(scale_factors, np.float32)
    s2 = _c(level_sigma2, np.float32)
    fv = FeatVec(len(node_ids), _ptr(node_ids), _ptr(offsets), _ptr(indices))
    kv = KeyFrameView(len(kps), _ptr(kps), _ptr(desc), _ptr(u_right), _ptr(hm), fv, _ptr(sf), _ptr(s2), len(sf))
    return Holder(kv, (kps, desc, u_right, hm, node_ids, offsets, indices, sf, s2))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.orbref_resize_linear.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci]
        L.orbref_gauss7.argtypes = [vp, ci, ci, ci, vp, ci]
        L.orbref_border101.argtypes = [vp, ci, ci, ci, vp, ci, ci]
        L.orbref_fast9.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, ci]
        L.orbref_fast_atan2.argtypes = [cf, cf]
        L.orbref_fast_atan2.restype = cf
        L.orbref_cv_round.argtypes = [cf]
        L.orbref_std_sort_perm.argtypes = [vp, vp, ci, vp]
        L.orbref_extractor_create.argtypes = [ci, cf, ci, ci, ci]
        L.orbref_extractor_create.restype = vp
        L.orbref_extractor_destroy.argtypes = [vp]
        L.orbref_extractor_tables.argtypes = [vp] * 7
        L.orbref_extract.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp]
        L.orbref_level_dims.argtypes = [vp, ci, vp, vp]
        for f in (L.orbref_level_image, L.orbref_level_bordered, L.orbref_level_blurred):
            f.argtypes = [vp, ci, vp]
            f.restype = vp
        L.orbref_level_candidates.argtypes = [vp, ci, vp, ci]
        L.orbref_level_keypoints.argtypes = [vp, ci, vp, ci]
        L.orbref_descriptor_distance.argtypes = [vp, vp]
        L.orbref_knn2.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp]
        L.orbref_distinctive_descriptor.argtypes = [vp, ci]
        L.orbref_distinctive_descriptor.restype = ci
        L.orbref_stereo_match.argtypes = [vp, vp, vp, vp, ci, vp, vp, ci, cf, cf, vp, vp]
        L.orbref_build_grid.argtypes = [vp, ci, cf, cf, cf, cf, vp, vp]
        L.orbref_features_in_area.argtypes = [vp, cf, cf, cf, ci, ci, vp]
        L.orbref_search_by_projection_map.argtypes = [vp, vp, cf, cf, ci, cf, vp]
        L.orbref_is_in_frustum.argtypes = [vp, vp, ci, cf, vp, vp, vp, vp, vp, vp, vp]
        L.orbref_is_in_frustum.restype = None
        L.orbref_search_by_projection_frame.argtypes = [vp, vp, ci, ci, vp]
        L.orbref_search_by_projection_frame_decisions.argtypes = [vp, vp, ci, vp, vp]
        L.orbref_search_for_triangulation.argtypes = [vp, vp, vp, cf, cf, ci, ci, ci, vp]
        L.orbref_search_by_bow.argtypes = [vp, vp, cf, ci, vp]
        L.orbref_search_for_triangulation_fisheye.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci, vp]
        L.orbref_triangulation_candidates.argtypes = [vp, vp, vp, vp, vp, ci]
        L.orbref_search_by_bow_kf.argtypes = [vp, vp, cf, ci, vp]
        L.orbref_search_by_bow_fisheye.argtypes = [vp, vp, ci, cf, ci, vp]
        L.orbref_bow_transform.argtypes = [vp, vp, ci, ci, vp, vp, vp]
        L.orbref_bow_transform.restype = None
        L.orbref_search_for_initialization.argtypes = [vp, vp, vp, ci, cf, ci, vp]
        L.orbref_remap_linear.argtypes = [vp, ci, ci, ci, vp, vp, ci, ci, vp, ci]
        L.orbref_remap_linear.restype = None
        L.orbref_cvt_gray.argtypes = [vp, ci, ci, ci, ci, ci, vp, ci]
        L.orbref_cvt_gray.restype = None
        L.orbref_fuse_match.argtypes = [vp, vp, vp, ci, vp, vp]
        L.orbref_fuse_match.restype = None
        L.orbref_extract_many.argtypes = [vp, ci, ci, ci, C.c_long, ci, cf, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp]
        L.orbref_stereo_many.argtypes = [vp, vp, ci, ci, ci, C.c_long, ci, cf, ci, ci, ci, cf, cf, ci, vp, vp, vp]
        _lib = L
    return _lib


# ---- primitives -------------------------------------------------------------------------------------------------
def primitives_kind():
    L = lib()
    L.orbref_primitives_kind.restype = C.c_char_p
    return L.orbref_primitives_kind().decode()


def resize_linear(src, dw, dh):
    src = _c(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orbref_resize_linear(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dw, dh, dw)
    return dst


def gauss7(src):
    src = _c(src, np.uint8)
    dst = np.empty_like(src)
    lib().orbref_gauss7(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dst.strides[0])
    return dst


def border101(src, border):
    src = _c(src, np.uint8)
    dst = np.empty((src.shape[0] + 2 * border, src.shape[1] + 2 * border), np.uint8)
    lib().orbref_border101(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dst.strides[0], border)
    return dst


def fast9(img, threshold):
    img = _c(img, np.uint8)
    cap = img.size
    xs, ys, sc = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orbref_fast9(_ptr(img), img.shape[1], img.shape[0], img.strides[0], threshold, _ptr(xs), _ptr(ys),
                           _ptr(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def fast_atan2(y, x):
    return lib().orbref_fast_atan2(float(y), float(x))


def std_sort_perm(key0, key1):
    key0 = _c(key0, np.int32)
    key1 = _c(key1, np.int32)
    perm = np.empty(len(key0), np.int32)
    lib().orbref_std_sort_perm(_ptr(key0), _ptr(key1), len(key0), _ptr(perm))
    return perm


# ---- extractor --------------------------------------------------------------------------------------------------
class Extractor:
    """Mirror of ORB_SLAM3::ORBextractor (include/ORBextractor.h:48-120) over the oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = lib().orbref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.scale = np.empty(nlevels, np.float32)
        self.inv_scale = np.empty(nlevels, np.float32)
        self.sigma2 = np.empty(nlevels, np.float32)
        self.inv_sigma2 = np.empty(nlevels, np.float32)
        self.features_per_level = np.empty(nlevels, np.int32)
        self.umax = np.empty(16, np.int32)
        lib().orbref_extractor_tables(self._h, _ptr(self.scale), _ptr(self.inv_scale), _ptr(self.sigma2),
                                      _ptr(self.inv_sigma2), _ptr(self.features_per_level), _ptr(self.umax))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orbref_extractor_destroy(self._h)
            self._h = None

    def __call__(self, img, lapping=(0, 0)):
        """Returns (mono_index, keypoints[KP_DTYPE], descriptors[n,32]); mono_index = -1 on empty input."""
        if img is None or img.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        img = _c(img, np.uint8)
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        rc = lib().orbref_extract(self._h, _ptr(img), img.shape[1], img.shape[0], img.strides[0], lapping[0],
                                  lapping[1], _ptr(kps), _ptr(desc), cap, C.byref(n), C.byref(mono))
        if rc != 0:
            raise RuntimeError("orbref_extract rc=%d n=%d" % (rc, n.value))
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def level_dims(self, level):
        w, h = C.c_int(0), C.c_int(0)
        lib().orbref_level_dims(self._h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def _view(self, fn, level, w, h):
        st = C.c_int(0)
        p = fn(self._h, level, C.byref(st))
        if not p:
            return None
        buf = (C.c_uint8 * (st.value * (h - 1) + w)).from_address(p)
        return np.lib.stride_tricks.as_strided(np.frombuffer(buf, np.uint8), (h, w), (st.value, 1)).copy()

    def level_image(self, level):
        w, h = self.level_dims(level)
        return self._view(lib().orbref_level_image, level, w, h)

    def level_bordered(self, level):
        w, h = self.level_dims(level)
        return self._view(lib().orbref_level_bordered, level, w + 38, h + 38)

    def level_blurred(self, level):
        w, h = self.level_dims(level)
        return self._view(lib().orbref_level_blurred, level, w, h)

    def _kps(self, fn, level):
        n = fn(self._h, level, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        fn(self._h, level, _ptr(out), n)
        return out[:n]

    def level_candidates(self, level):
        return self._kps(lib().orbref_level_candidates, level)

    def level_keypoints(self, level):
        return self._kps(lib().orbref_level_keypoints, level)


# ---- matching ---------------------------------------------------------------------------------------------------
def descriptor_distance(a, b):
    a = _c(a, np.uint8)
    b = _c(b, np.uint8)
    return lib().orbref_descriptor_distance(_ptr(a), _ptr(b))


def knn2(q, t):
    q = _c(q, np.uint8).reshape(-1, 32)
    t = _c(t, np.uint8).reshape(-1, 32)
    out = [np.empty(len(q), np.int32) for _ in range(4)]
    lib().orbref_knn2(_ptr(q), len(q), _ptr(t), len(t), *[_ptr(o) for o in out])
    return tuple(out)  # idx1, d1, idx2, d2


def distinctive_descriptor(desc):
    """MapPoint::ComputeDistinctiveDescriptors on one point's observed descriptors [n, 32] -> index or -1."""
    desc = _c(desc, np.uint8).reshape(-1, 32)
    return int(lib().orbref_distinctive_descriptor(_ptr(desc), len(desc)))


def stereo_match(ex_l, ex_r, kps_l, desc_l, kps_r, desc_r, mbf, mb):
    kps_l, kps_r = _c(kps_l, KP_DTYPE), _c(kps_r, KP_DTYPE)
    desc_l, desc_r = _c(desc_l, np.uint8), _c(desc_r, np.uint8)
    ur = np.empty(len(kps_l), np.float32)
    dp = np.empty(len(kps_l), np.float32)
    n = lib().orbref_stereo_match(ex_l._h, ex_r._h, _ptr(kps_l), _ptr(desc_l), len(kps_l), _ptr(kps_r), _ptr(desc_r),
                                  len(kps_r), mbf, mb, _ptr(ur), _ptr(dp))
    return n, ur, dp


def build_grid(kps, min_x, min_y, inv_w, inv_h):
    kps = _c(kps, KP_DTYPE)
    off = np.empty(GRID_COLS * GRID_ROWS + 1, np.int32)
    items = np.empty(max(len(kps), 1), np.int32)
    lib().orbref_build_grid(_ptr(kps), len(kps), min_x, min_y, inv_w, inv_h, _ptr(off), _ptr(items))
    return off, items[:off[-1]].copy()


def features_in_area(fv, x, y, r, min_level, max_level):
    out = np.empty(max(fv.struct.n, 1), np.int32)
    n = lib().orbref_features_in_area(fv.ref(), x, y, r, min_level, max_level, _ptr(out))
    return out[:n].copy()


def search_by_projection_map(fv, mps, th, nnratio, far_points=False, th_far=0.0):
    assign = np.empty(max(fv.struct.n, 1), np.int32)
    n = lib().orbref_search_by_projection_map(fv.ref(), mps.ref(), th, nnratio, int(far_points), th_far, _ptr(assign))
    return n, assign[:fv.struct.n]


def search_by_projection_frame_decisions(fv, pts, max_dist=100, fn=None):
    """One camera's candidate loop of SearchByProjection(CurrentFrame, LastFrame): (decisions[m], window[m])."""
    m = pts.struct.m
    dec, win = np.empty(max(m, 1), np.int32), np.empty(max(m, 1), np.int32)
    fn = fn or lib().orbref_search_by_projection_frame_decisions
    fn.restype = C.c_int
    fn(fv.ref(), pts.ref(), C.c_int(max_dist), _ptr(dec), _ptr(win))
    return dec[:m], win[:m]


def search_by_projection_frame(fv, pts, max_dist=100, check_orientation=True):
    assign = np.empty(max(fv.struct.n, 1), np.int32)
    n = lib().orbref_search_by_projection_frame(fv.ref(), pts.ref(), max_dist, int(check_orientation), _ptr(assign))
    return n, assign[:fv.struct.n]


def search_for_triangulation(kf1, kf2, F12, ep, only_stereo=False, coarse=False, check_orientation=True):
    F12 = _c(F12, np.float32).reshape(9)
    m = np.empty(max(kf1.struct.n, 1), np.int32)
    n = lib().orbref_search_for_triangulation(kf1.ref(), kf2.ref(), _ptr(F12), float(ep[0]), float(ep[1]),
                                              int(only_stereo), int(coarse), int(check_orientation), _ptr(m))
    return n, m[:kf1.struct.n]


def search_for_triangulation_fisheye(kf1, n_left1, kf2, n_left2, F12x4, only_stereo=False, coarse=False,
                                     check_orientation=True, fn=None):
    """SearchForTriangulation on two-camera KeyFrames; F12x4[2 * right1 + right2] = the selected pair's matrix."""
    F = np.ascontiguousarray(F12x4, np.float32).reshape(36)
    m = np.empty(max(kf1.struct.n, 1), np.int32)
    n = (fn or lib().orbref_search_for_triangulation_fisheye)(kf1.ref(), int(n_left1), kf2.ref(), int(n_left2), _ptr(F),
                                                              int(only_stereo), int(coarse), int(check_orientation),
                                                              _ptr(m))
    return n, m[:kf1.struct.n]


def triangulation_candidates(kf1, kf2):
    """(offsets[n1 + 1], idx2[total], dist[total]): the descriptor part of SearchForTriangulation, :973-988."""
    n1 = kf1.struct.n
    off = np.zeros(n1 + 1, np.int32)
    total = lib().orbref_triangulation_candidates(kf1.ref(), kf2.ref(), _ptr(off), None, None, 0)
    idx, dist = np.empty(max(total, 1), np.int32), np.empty(max(total, 1), np.int32)
    lib().orbref_triangulation_candidates(kf1.ref(), kf2.ref(), _ptr(off), _ptr(idx), _ptr(dist), total)
    return off, idx[:total], dist[:total]


def search_by_bow(kf, frame, nnratio=0.7, check_orientation=True):
    m = np.empty(max(frame.struct.n, 1), np.int32)
    n = lib().orbref_search_by_bow(kf.ref(), frame.ref(), float(nnratio), int(check_orientation), _ptr(m))
    return n, m[:frame.struct.n]


def search_by_bow_fisheye(kf, frame, n_left_f, nnratio=0.7, check_orientation=True):
    """SearchByBoW(KeyFrame*, Frame&, ...) on a two-camera Frame: rows [0, n_left_f) of the frame view are the left
    camera's, the rest the right camera's (src/ORBmatcher.cc:274-365)."""
    m = np.empty(max(frame.struct.n, 1), np.int32)
    n = lib().orbref_search_by_bow_fisheye(kf.ref(), frame.ref(), int(n_left_f), float(nnratio), int(check_orientation),
                                           _ptr(m))
    return n, m[:frame.struct.n]


def search_by_bow_kf(kf1, kf2, nnratio=0.8, check_orientation=True):
    m = np.empty(max(kf1.struct.n, 1), np.int32)
    n = lib().orbref_search_by_bow_kf(kf1.ref(), kf2.ref(), float(nnratio), int(check_orientation), _ptr(m))
    return n, m[:kf1.struct.n]


def make_vocabulary(depth, child_offsets, children, descriptors, word_id, weight):
    arrs = (_c(child_offsets, np.int32), _c(children, np.uint32), _c(descriptors, np.uint8), _c(word_id, np.uint32),
            _c(weight, np.float64))
    return Holder(Vocabulary(len(arrs[0]) - 1, int(depth), *[_ptr(a) for a in arrs]), arrs)


def bow_transform(voc, desc, levelsup=4):
    desc = _c(desc, np.uint8).reshape(-1, 32)
    n = len(desc)
    w, wt, nd = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.float64), np.empty(max(n, 1), np.uint32)
    lib().orbref_bow_transform(voc.ref(), _ptr(desc), n, int(levelsup), _ptr(w), _ptr(wt), _ptr(nd))
    return w[:n], wt[:n], nd[:n]


def search_for_initialization(f1, f2, prev_xy, window_size=100, nnratio=0.9, check_orientation=True):
    prev = _c(prev_xy, np.float32).reshape(-1, 2)
    m = np.empty(max(f1.struct.n, 1), np.int32)
    n = lib().orbref_search_for_initialization(f1.ref(), f2.ref(), _ptr(prev), int(window_size), float(nnratio),
                                               int(check_orientation), _ptr(m))
    return n, m[:f1.struct.n]


def remap_linear(img, mapx, mapy):
    """cv::remap(img, mapx, mapy, INTER_LINEAR) for a uint8 [h, w] image and float32 maps [dh, dw]."""
    img = _c(img, np.uint8)
    mapx, mapy = _c(mapx, np.float32), _c(mapy, np.float32)
    dh, dw = mapx.shape
    out = np.empty((dh, dw), np.uint8)
    lib().orbref_remap_linear(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(mapx), _ptr(mapy), dw, dh,
                              _ptr(out), dw)
    return out


def cvt_gray(img, rgb=False):
    """cv::cvtColor(img, COLOR_{BGR,RGB,BGRA,RGBA}2GRAY) for an [h, w, 3|4] uint8 image."""
    img = _c(img, np.uint8)
    h, w, c = img.shape
    out = np.empty((h, w), np.uint8)
    lib().orbref_cvt_gray(_ptr(img), w, h, img.strides[0], c, int(rgb), _ptr(out), w)
    return out


def fuse_match(kf, inv_level_sigma2, pts, chi2_gate=True):
    inv = _c(inv_level_sigma2, np.float32)
    m = pts.struct.m
    bi, bd = np.empty(max(m, 1), np.int32), np.empty(max(m, 1), np.int32)
    lib().orbref_fuse_match(kf.ref(), _ptr(inv), pts.ref(), int(chi2_gate), _ptr(bi), _ptr(bd))
    return bi[:m], bd[:m]


def extract_many(imgs, nfeatures, scale_factor, nlevels, ini_th, min_th, lapping, threads):
    imgs = _c(imgs, np.uint8)
    n, h, w = imgs.shape
    counts = np.zeros(n, np.int32)
    cap = nfeatures + 8 * nlevels + 64
    lib().orbref_extract_many(_ptr(imgs), n, w, h, w * h, nfeatures, scale_factor, nlevels, ini_th, min_th,
                              lapping[0], lapping[1], threads, None, None, cap, _ptr(counts))
    return counts


def stereo_many(imgs_l, imgs_r, nfeatures, scale_factor, nlevels, ini_th, min_th, mbf, mb, threads):
    imgs_l, imgs_r = _c(imgs_l, np.uint8), _c(imgs_r, np.uint8)
    n, h, w = imgs_l.shape
    cl, cr, mt = (np.zeros(n, np.int32) for _ in range(3))
    lib().orbref_stereo_many(_ptr(imgs_l), _ptr(imgs_r), n, w, h, w * h, nfeatures, scale_factor, nlevels, ini_th,
                             min_th, mbf, mb, threads, _ptr(cl), _ptr(cr), _ptr(mt))
    return cl, cr, mt
