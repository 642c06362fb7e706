"""ctypes mirrors of the flat views in include/orbx_types.h (what the C++ shim builds from Frame / MapPoint / KeyFrame).

Each `make_*` returns a Holder that keeps the numpy arrays alive next to the struct that points into them.
"""
import ctypes as C

import numpy as np

from .lib import KP_DTYPE

GRID_COLS, GRID_ROWS = 64, 48  # FRAME_GRID_COLS / FRAME_GRID_ROWS, include/Frame.h:41-42


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


class Vocabulary(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("depth", C.c_int32), ("child_offsets", C.c_void_p), ("children", C.c_void_p),
                ("descriptors", C.c_void_p), ("word_id", C.c_void_p), ("weight", C.c_void_p)]


class FeatVec(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_ids", C.c_void_p), ("offsets", C.c_void_p), ("indices", C.c_void_p)]


class KeyFrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("kps", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p),
                ("has_mappoint", C.c_void_p), ("featvec", FeatVec), ("scale_factors", C.c_void_p),
                ("level_sigma2", C.c_void_p), ("n_levels", C.c_int32)]


class FisheyeView(C.Structure):
    _fields_ = [("n_left", C.c_int32), ("n_right", C.c_int32), ("kps_left", C.c_void_p), ("kps_right", C.c_void_p),
                ("desc", C.c_void_p), ("occupied", C.c_void_p), ("grid_left", Grid), ("grid_right", Grid),
                ("left_to_right", C.c_void_p), ("right_to_left", C.c_void_p), ("scale_factors", C.c_void_p),
                ("n_levels", C.c_int32)]


class MapPointsRight(C.Structure):
    _fields_ = [("track_in_view_r", C.c_void_p), ("proj_x_r", C.c_void_p), ("proj_y_r", C.c_void_p),
                ("level_r", C.c_void_p), ("view_cos_r", C.c_void_p)]


class Frustum(C.Structure):
    """orbx_frustum (include/orbx_types.h): what Frame::isInFrustum reads from the Frame; 104 bytes."""
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("fx", C.c_float),
                ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float), ("min_x", C.c_float),
                ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float), ("log_scale_factor", C.c_float),
                ("n_levels", C.c_int32)]


FRUSTUM_DTYPE = np.dtype([("Rcw", "<f4", 9), ("tcw", "<f4", 3), ("Ow", "<f4", 3), ("fx", "<f4"), ("fy", "<f4"),
                          ("cx", "<f4"), ("cy", "<f4"), ("mbf", "<f4"), ("min_x", "<f4"), ("max_x", "<f4"),
                          ("min_y", "<f4"), ("max_y", "<f4"), ("log_scale_factor", "<f4"), ("n_levels", "<i4")])
assert FRUSTUM_DTYPE.itemsize == 104 and C.sizeof(Frustum) == 104


class LocalMap(C.Structure):
    _fields_ = [("m", C.c_int32), ("n_maps", C.c_int32), ("pos", C.c_void_p), ("normal", C.c_void_p),
                ("min_dist", C.c_void_p), ("max_dist", C.c_void_p), ("skip", C.c_void_p), ("has_obs", C.c_void_p),
                ("desc", C.c_void_p)]


class TrackParams(C.Structure):
    _fields_ = [("viewing_cos_limit", C.c_float), ("th", C.c_float), ("nnratio", C.c_float), ("far_points", C.c_int32),
                ("th_far", C.c_float), ("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float),
                ("inv_h", C.c_float), ("cand_per_frame", C.c_int32)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Holder:
    def __init__(self, struct, keep):
        self.struct, self.keep = struct, keep

    def ref(self):
        return C.byref(self.struct)


def assign_features_to_grid(kps, min_x, min_y, inv_w, inv_h):
    """Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:520-547, 833-844) as a CSR — host-side, like in the
    reference, where the grid is built by the Frame constructor (the caller of the hot path)."""
    px = np.round((kps["x"].astype(np.float32) - np.float32(min_x)) * np.float32(inv_w))
    py = np.round((kps["y"].astype(np.float32) - np.float32(min_y)) * np.float32(inv_h))
    # np.round is half-to-even, C round() is half-away-from-zero: fix the exact .5 cases
    fx = (kps["x"].astype(np.float32) - np.float32(min_x)) * np.float32(inv_w)
    fy = (kps["y"].astype(np.float32) - np.float32(min_y)) * np.float32(inv_h)
    px = np.where(np.abs(fx - np.trunc(fx)) == 0.5, np.trunc(fx) + np.sign(fx), px).astype(np.int64)
    py = np.where(np.abs(fy - np.trunc(fy)) == 0.5, np.trunc(fy) + np.sign(fy), py).astype(np.int64)
    ok = (px >= 0) & (px < GRID_COLS) & (py >= 0) & (py < GRID_ROWS)
    cell = np.where(ok, px * GRID_ROWS + py, GRID_COLS * GRID_ROWS)
    order = np.argsort(cell, kind="stable")
    counts = np.bincount(cell, minlength=GRID_COLS * GRID_ROWS + 1)[:GRID_COLS * GRID_ROWS]
    offsets = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
    offsets[1:] = np.cumsum(counts)
    items = order[:offsets[-1]].astype(np.int32)
    return offsets, items


def make_frame_view(kps, desc, u_right, occupied, offsets, items, min_x, min_y, inv_w, inv_h, scale_factors):
    kps = _c(kps, KP_DTYPE)
    desc = _c(desc, np.uint8)
    u_right = None if u_right is None else _c(u_right, np.float32)
    occupied = _c(occupied, np.uint8)
    offsets, items = _c(offsets, np.int32), _c(items, np.int32)
    sf = _c(scale_factors, np.float32)
    g = Grid(_p(offsets), _p(items), float(min_x), float(min_y), float(inv_w), float(inv_h))
    fv = FrameView(len(kps), _p(kps), _p(desc), _p(u_right), _p(occupied), g, _p(sf), len(sf))
    return Holder(fv, (kps, desc, u_right, occupied, offsets, items, sf))


def make_mappoints(track_in_view, proj_x, proj_y, proj_xr, level, view_cos, depth, has_obs, desc):
    arrs = (_c(track_in_view, np.uint8), _c(proj_x, np.float32), _c(proj_y, np.float32), _c(proj_xr, np.float32),
            _c(level, np.int32), _c(view_cos, np.float32), _c(depth, np.float32), _c(has_obs, np.uint8),
            _c(desc, np.uint8))
    return Holder(MapPoints(len(arrs[0]), *[_p(a) for a in arrs]), arrs)


def make_projected(u, v, u_right, radius, min_level, max_level, angle, has_obs, desc):
    arrs = (_c(u, np.float32), _c(v, np.float32), None if u_right is None else _c(u_right, np.float32),
            _c(radius, np.float32), _c(min_level, np.int32), _c(max_level, np.int32), _c(angle, np.float32),
            _c(has_obs, np.uint8), _c(desc, np.uint8))
    return Holder(Projected(len(arrs[0]), *[_p(a) for a in arrs]), arrs)


def make_vocabulary(depth, child_offsets, children, descriptors, word_id, weight):
    """orbx_vocabulary: DBoW2's m_nodes flattened (include/orbx_types.h)."""
    arrs = (_c(child_offsets, np.int32), _c(children, np.uint32), _c(descriptors, np.uint8), _c(word_id, np.uint32),
            _c(weight, np.float64))
    return Holder(Vocabulary(len(arrs[0]) - 1, int(depth), *[_p(a) for a in arrs]), arrs)


def assemble_bow(word_id, weight, node_id):
    """What the shim does with the per-feature answers (TemplatedVocabulary.h:1147-1160, 1198; BowVector.cpp:34-84):
    BowVector = {word: sum of weights in feature order}, L1-normalised in ascending word order; FeatureVector =
    {node: feature indices in order}. Plain Python floats are the reference's doubles."""
    bow, fv = {}, {}
    for i, (w, wt, nd) in enumerate(zip(word_id.tolist(), weight.tolist(), node_id.tolist())):
        if wt > 0:
            bow[w] = bow[w] + wt if w in bow else wt
            fv.setdefault(nd, []).append(i)
    norm = 0.0
    for w in sorted(bow):
        norm += abs(bow[w])
    if norm > 0.0:
        for w in bow:
            bow[w] /= norm
    return bow, fv


def make_keyframe_view(kps, desc, u_right, has_mappoint, node_ids, offsets, indices, scale_factors, level_sigma2):
    kps = _c(kps, KP_DTYPE)
    desc = _c(desc, np.uint8)
    u_right = None if u_right is None else _c(u_right, np.float32)
    hm = _c(has_mappoint, np.uint8)
    node_ids, offsets, indices = _c(node_ids, np.uint32), _c(offsets, np.int32), _c(indices, np.uint32)
    sf, s2 = _c(scale_factors, np.float32), _c(level_sigma2, np.float32)
    fv = FeatVec(len(node_ids), _p(node_ids), _p(offsets), _p(indices))
    kv = KeyFrameView(len(kps), _p(kps), _p(desc), _p(u_right), _p(hm), fv, _p(sf), _p(s2), len(sf))
    return Holder(kv, (kps, desc, u_right, hm, node_ids, offsets, indices, sf, s2))


def make_local_map(pos, normal, min_dist, max_dist, skip, has_obs, desc):
    """orbx_local_map over host arrays: pos / normal [n_maps, m, 3], the others [n_maps, m] (desc [n_maps, m, 32]); a
    single map may omit the leading axis."""
    pos = _c(pos, np.float32)
    if pos.ndim == 2:
        pos = pos[None]
    n_maps, m = pos.shape[0], pos.shape[1]
    arrs = (pos, _c(normal, np.float32).reshape(n_maps, m, 3), _c(min_dist, np.float32).reshape(n_maps, m),
            _c(max_dist, np.float32).reshape(n_maps, m), None if skip is None else _c(skip, np.uint8).reshape(n_maps, m),
            _c(has_obs, np.uint8).reshape(n_maps, m), _c(desc, np.uint8).reshape(n_maps, m, 32))
    return Holder(LocalMap(m, n_maps, *[_p(a) for a in arrs]), arrs)


def make_local_map_device(m, n_maps, pos, normal, min_dist, max_dist, skip, has_obs, desc):
    """orbx_local_map whose array pointers are device addresses (integers, e.g. torch data_ptr())."""
    vp = C.c_void_p
    return Holder(LocalMap(m, n_maps, vp(pos), vp(normal), vp(min_dist), vp(max_dist), vp(skip) if skip else None,
                           vp(has_obs), vp(desc)), ())


def make_track_params(width, height, th=1.0, nnratio=0.8, viewing_cos_limit=0.5, far_points=False, th_far=0.0,
                      min_x=0.0, min_y=0.0, max_x=None, max_y=None, cand_per_frame=0):
    """Tracking::SearchLocalPoints' scalars; the grid cell sizes as Frame computes them (src/Frame.cc:230-233)."""
    max_x = float(width) if max_x is None else max_x
    max_y = float(height) if max_y is None else max_y
    inv_w = np.float32(GRID_COLS) / (np.float32(max_x) - np.float32(min_x))
    inv_h = np.float32(GRID_ROWS) / (np.float32(max_y) - np.float32(min_y))
    return TrackParams(viewing_cos_limit, th, nnratio, int(far_points), th_far, min_x, min_y, float(inv_w), float(inv_h),
                       cand_per_frame)


def make_fisheye_view(kps_left, kps_right, desc, occupied, min_x, min_y, inv_w, inv_h, left_to_right, right_to_left,
                      scale_factors):
    """orbx_fisheye_view of a two-camera Frame (Nleft != -1); both grids are built here (Frame::AssignFeaturesToGrid)."""
    kl, kr = _c(kps_left, KP_DTYPE), _c(kps_right, KP_DTYPE)
    desc, occupied = _c(desc, np.uint8), _c(occupied, np.uint8)
    l2r, r2l = _c(left_to_right, np.int32), _c(right_to_left, np.int32)
    sf = _c(scale_factors, np.float32)
    ol, il = assign_features_to_grid(kl, min_x, min_y, inv_w, inv_h)
    o_r, ir = assign_features_to_grid(kr, min_x, min_y, inv_w, inv_h)
    ol, il, o_r, ir = _c(ol, np.int32), _c(il, np.int32), _c(o_r, np.int32), _c(ir, np.int32)
    gl = Grid(_p(ol), _p(il), float(min_x), float(min_y), float(inv_w), float(inv_h))
    gr = Grid(_p(o_r), _p(ir), float(min_x), float(min_y), float(inv_w), float(inv_h))
    v = FisheyeView(len(kl), len(kr), _p(kl), _p(kr), _p(desc), _p(occupied), gl, gr, _p(l2r), _p(r2l), _p(sf), len(sf))
    return Holder(v, (kl, kr, desc, occupied, ol, il, o_r, ir, l2r, r2l, sf))


def make_mappoints_right(track_in_view_r, proj_x_r, proj_y_r, level_r, view_cos_r):
    arrs = (_c(track_in_view_r, np.uint8), _c(proj_x_r, np.float32), _c(proj_y_r, np.float32), _c(level_r, np.int32),
            _c(view_cos_r, np.float32))
    return Holder(MapPointsRight(*[_p(a) for a in arrs]), arrs)
