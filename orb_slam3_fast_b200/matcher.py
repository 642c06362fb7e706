"""Host-side mirror of ORB_SLAM3::ORBmatcher (include/ORBmatcher.h:38-133) and of the two Frame methods on the hot
path (ComputeStereoMatches, the knnMatch of ComputeStereoFishEyeMatches) over the orbm C ABI.
"""
import ctypes as C

import numpy as np

from . import lib as _l
from .lib import KP_DTYPE, OrbxError

TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30  # src/ORBmatcher.cc:35-37


def _bind(L):
    if getattr(L, "_orbm_bound", False):
        return
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.orbm_create.argtypes = [C.POINTER(vp), ci]
    L.orbm_destroy.argtypes = [vp]
    L.orbm_destroy.restype = None
    L.orbm_last_error.argtypes = [vp]
    L.orbm_last_error.restype = C.c_char_p
    L.orbm_descriptor_distance_batch.argtypes = [vp, vp, vp, ci, vp]
    L.orbm_knn2.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp, vp]
    L.orbm_distinctive_descriptors.argtypes = [vp, vp, vp, ci, vp]
    L.orbm_knn2_device.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp, vp, vp]
    L.orbm_stereo_match.argtypes = [vp, vp, vp, ci, vp, vp, ci, vp, vp, ci, cf, cf, vp, vp, vp]
    L.orbm_stereo_match_batch_device.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, ci, cf, cf, vp, vp, vp, vp]
    L.orbm_stereo_frames_batch.argtypes = [vp, vp, vp, ci, vp, vp, ci, ci, ci, C.c_int64, cf, cf, vp, vp, vp, vp, vp, vp,
                                           ci, vp, vp, vp]
    L.orbm_search_by_projection_map.argtypes = [vp, vp, vp, cf, cf, ci, cf, vp, vp]
    L.orbm_search_by_projection_frame.argtypes = [vp, vp, vp, ci, ci, vp, vp]
    L.orbm_search_by_projection_frame_decisions.argtypes = [vp, vp, vp, ci, vp, vp]
    L.orbm_assign_features_to_grid.argtypes = [vp, vp, ci, cf, cf, cf, cf, vp, vp]
    L.orbm_search_by_projection_map_resident.argtypes = [vp, vp, ci, ci, vp, vp, cf, cf, cf, cf, vp, cf, cf, ci, cf, vp,
                                                         vp]
    L.orbm_search_by_projection_frame_resident.argtypes = [vp, vp, ci, ci, vp, vp, cf, cf, cf, cf, vp, ci, ci, vp, vp]
    L.orbm_search_for_triangulation.argtypes = [vp, vp, vp, vp, cf, cf, ci, ci, ci, vp, vp]
    L.orbm_search_by_bow.argtypes = [vp, vp, vp, cf, ci, vp, vp]
    L.orbm_fuse_match.argtypes = [vp, vp, vp, vp, ci, vp, vp]
    L.orbm_set_vocabulary.argtypes = [vp, vp]
    L.orbm_bow_transform.argtypes = [vp, vp, ci, ci, vp, vp, vp]
    L.orbm_search_for_initialization.argtypes = [vp, vp, vp, vp, ci, cf, ci, vp, vp]
    L.orbm_search_by_bow_kf.argtypes = [vp, vp, vp, cf, ci, vp, vp]
    L.orbm_search_by_bow_fisheye.argtypes = [vp, vp, vp, ci, cf, ci, vp, vp]
    L.orbm_triangulation_candidates.argtypes = [vp, vp, vp, vp, vp, vp, ci, vp]
    L.orbm_is_in_frustum.argtypes = [vp, vp, vp, ci, cf, vp, vp, vp, vp, vp, vp, vp, vp]
    L.orbm_track_local_map_batch_device.argtypes = [vp, vp, ci, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                                    vp]
    L.orbm_stereo_track_frames_batch.argtypes = [vp, vp, vp, ci, vp, vp, ci, ci, ci, C.c_int64, cf, cf, vp, vp, vp, vp,
                                                 vp, vp, vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp]
    L.orbm_stereo_track_frames_batch_multi.argtypes = [ci, vp, vp, vp, ci, vp, vp, ci, ci, ci, C.c_int64, cf, cf, vp, vp, vp,
                                                       vp, vp, vp, vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp]
    L.orbm_search_by_projection_map_fisheye.argtypes = [vp, vp, vp, vp, cf, cf, ci, cf, vp, vp]
    L._orbm_bound = True


class ORBmatcher:
    """ORBmatcher(nnratio = 0.6, checkOri = true) (include/ORBmatcher.h:38)."""

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self._L = _l.lib()
        _bind(self._L)
        h = C.c_void_p()
        rc = self._L.orbm_create(C.byref(h), device)
        if rc != 0:
            raise OrbxError(rc, self._L.orbm_last_error(None).decode())
        self._h = h
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbm_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise OrbxError(rc, self._L.orbm_last_error(self._h).decode())
        return rc

    # static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) — src/ORBmatcher.cc:1959-1973
    def DescriptorDistanceBatch(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.empty(len(a), np.int32)
        self._check(self._L.orbm_descriptor_distance_batch(self._h, _l.ptr(a), _l.ptr(b), len(a), _l.ptr(out)))
        return out

    # cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, matches, 2) — src/Frame.cc:1293
    # void MapPoint::ComputeDistinctiveDescriptors() — src/MapPoint.cc:372-441, batched: lists[p] = [n_p, 32] arrays
    def ComputeDistinctiveDescriptors(self, lists):
        offsets = np.zeros(len(lists) + 1, np.int32)
        offsets[1:] = np.cumsum([len(x) for x in lists])
        desc = (np.concatenate([np.asarray(x, np.uint8).reshape(-1, 32) for x in lists]) if len(lists) and offsets[-1]
                else np.zeros((1, 32), np.uint8))
        desc = np.ascontiguousarray(desc)
        best = np.empty(max(len(lists), 1), np.int32)
        self._check(self._L.orbm_distinctive_descriptors(self._h, _l.ptr(desc), _l.ptr(offsets), len(lists),
                                                         _l.ptr(best)))
        return best[:len(lists)]

    def knnMatch2(self, q, t):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        out = [np.empty(len(q), np.int32) for _ in range(4)]
        self._check(self._L.orbm_knn2(self._h, _l.ptr(q), len(q), _l.ptr(t), len(t), *[_l.ptr(o) for o in out]))
        return tuple(out)  # idx1, d1, idx2, d2

    def knnMatch2_device(self, d_q, nq, d_t, nt, d_idx1, d_d1, d_idx2, d_d2, stream=0):
        self._check(self._L.orbm_knn2_device(self._h, C.c_void_p(d_q), nq, C.c_void_p(d_t), nt, C.c_void_p(d_idx1),
                                             C.c_void_p(d_d1), C.c_void_p(d_idx2), C.c_void_p(d_d2),
                                             C.c_void_p(stream) if stream else None))

    # void Frame::ComputeStereoMatches() — src/Frame.cc:921-1084
    def ComputeStereoMatches(self, ex_left, ex_right, kps_l, desc_l, kps_r, desc_r, mbf, mb, frame=0):
        kps_l, kps_r = np.ascontiguousarray(kps_l, KP_DTYPE), np.ascontiguousarray(kps_r, KP_DTYPE)
        desc_l, desc_r = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
        ur = np.empty(len(kps_l), np.float32)
        dp = np.empty(len(kps_l), np.float32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_stereo_match(self._h, ex_left._h, ex_right._h, frame, _l.ptr(kps_l), _l.ptr(desc_l),
                                              len(kps_l), _l.ptr(kps_r), _l.ptr(desc_r), len(kps_r), mbf, mb,
                                              _l.ptr(ur), _l.ptr(dp), C.byref(nm)))
        return nm.value, ur, dp

    def ComputeStereoMatches_device(self, ex_left, ex_right, n_pairs, d_kps_l, d_desc_l, d_n_l, d_kps_r, d_desc_r,
                                    d_n_r, cap, mbf, mb, d_u_right, d_depth, d_n_matched, stream=0):
        vp = C.c_void_p
        self._check(self._L.orbm_stereo_match_batch_device(self._h, ex_left._h, ex_right._h, n_pairs, vp(d_kps_l),
                                                           vp(d_desc_l), vp(d_n_l), vp(d_kps_r), vp(d_desc_r),
                                                           vp(d_n_r), cap, mbf, mb, vp(d_u_right), vp(d_depth),
                                                           vp(d_n_matched), vp(stream) if stream else None))

    # Frame::Frame(stereo) hot path, batched, host buffers: extract x2 + ComputeStereoMatches — src/Frame.cc:149-279
    @staticmethod
    def alloc_stereo_outputs(n_pairs, cap, empty=np.empty):
        """Output arrays of StereoFramesBatch; pass an `empty` that returns pinned memory for asynchronous copies."""
        return dict(kps_l=empty((n_pairs, cap), KP_DTYPE), desc_l=empty((n_pairs, cap, 32), np.uint8),
                    n_l=empty((n_pairs,), np.int32), kps_r=empty((n_pairs, cap), KP_DTYPE),
                    desc_r=empty((n_pairs, cap, 32), np.uint8), n_r=empty((n_pairs,), np.int32),
                    u_right=empty((n_pairs, cap), np.float32), depth=empty((n_pairs, cap), np.float32),
                    n_matched=empty((n_pairs,), np.int32))

    def StereoFramesBatch(self, ex_left, ex_right, imgs_l, imgs_r, mbf, mb, out=None):
        """imgs_l / imgs_r: uint8 [n, h, w] host arrays (row stride may exceed w). Returns the dict of outputs."""
        assert imgs_l.shape == imgs_r.shape and imgs_l.strides == imgs_r.strides and imgs_l.strides[2] == 1
        n, h, w = imgs_l.shape
        cap = ex_left.capacity
        if out is None:
            out = self.alloc_stereo_outputs(n, cap)
        self._check(self._L.orbm_stereo_frames_batch(
            self._h, ex_left._h, ex_right._h, n, _l.ptr(imgs_l), _l.ptr(imgs_r), w, h, imgs_l.strides[1],
            imgs_l.strides[0], mbf, mb, _l.ptr(out["kps_l"]), _l.ptr(out["desc_l"]), _l.ptr(out["n_l"]),
            _l.ptr(out["kps_r"]), _l.ptr(out["desc_r"]), _l.ptr(out["n_r"]), cap, _l.ptr(out["u_right"]),
            _l.ptr(out["depth"]), _l.ptr(out["n_matched"])))
        return out

    # bool Frame::isInFrustum(MapPoint*, viewingCosLimit) over a local map — src/Frame.cc:632-699, Tracking.cc:3288-3300
    def IsInFrustum(self, frustum, local_map, map_index=0, viewingCosLimit=0.5, out=None):
        """frustum: one views.FRUSTUM_DTYPE record; local_map: views.make_local_map(...). Returns (n_in_view, dict) with
        the orbx_mappoints arrays track_in_view, proj_x, proj_y, proj_xr, level, view_cos, depth."""
        m = local_map.struct.m
        fr = np.ascontiguousarray(frustum).reshape(1)
        if out is None:
            out = dict(track_in_view=np.zeros(m, np.uint8), proj_x=np.zeros(m, np.float32),
                       proj_y=np.zeros(m, np.float32), proj_xr=np.zeros(m, np.float32), level=np.zeros(m, np.int32),
                       view_cos=np.zeros(m, np.float32), depth=np.zeros(m, np.float32))
        nv = C.c_int32(0)
        self._check(self._L.orbm_is_in_frustum(self._h, _l.ptr(fr), local_map.ref(), map_index, viewingCosLimit,
                                               _l.ptr(out["track_in_view"]), _l.ptr(out["proj_x"]),
                                               _l.ptr(out["proj_y"]), _l.ptr(out["proj_xr"]), _l.ptr(out["level"]),
                                               _l.ptr(out["view_cos"]), _l.ptr(out["depth"]), C.byref(nv)))
        return nv.value, out

    # void Tracking::SearchLocalPoints() for the frames of one device-resident extract batch — src/Tracking.cc:3249-3330
    def TrackLocalMapBatch_device(self, extractor, n_frames, d_kps, d_desc, d_n, cap, d_u_right, d_occupied, d_frustums,
                                  local_map_device, d_map_index, params, d_assign, d_nmatches, d_n_in_view, d_status,
                                  stream=0):
        vp = C.c_void_p
        self._check(self._L.orbm_track_local_map_batch_device(
            self._h, extractor._h, n_frames, vp(d_kps), vp(d_desc), vp(d_n), cap, vp(d_u_right) if d_u_right else None,
            vp(d_occupied) if d_occupied else None, vp(d_frustums), local_map_device.ref(),
            vp(d_map_index) if d_map_index else None, C.byref(params), vp(d_assign), vp(d_nmatches), vp(d_n_in_view),
            vp(d_status), vp(stream) if stream else None))

    # the stereo Frame constructor's hot path + Tracking::SearchLocalPoints per pair, host buffers (configs[3] end to end)
    @staticmethod
    def alloc_track_outputs(n_pairs, cap, empty=np.empty):
        out = ORBmatcher.alloc_stereo_outputs(n_pairs, cap, empty)
        out.update(assign=empty((n_pairs, cap), np.int32), nmatches=empty((n_pairs,), np.int32),
                   n_in_view=empty((n_pairs,), np.int32))
        return out

    def StereoTrackFramesBatch(self, ex_left, ex_right, imgs_l, imgs_r, mbf, mb, frustums, local_map, params,
                               map_index=None, occupied=None, out=None):
        assert imgs_l.shape == imgs_r.shape and imgs_l.strides == imgs_r.strides and imgs_l.strides[2] == 1
        n, h, w = imgs_l.shape
        cap = ex_left.capacity
        if out is None:
            out = self.alloc_track_outputs(n, cap)
        assert frustums.dtype.itemsize == 104 and len(frustums) == n and frustums.flags.c_contiguous
        self._check(self._L.orbm_stereo_track_frames_batch(
            self._h, ex_left._h, ex_right._h, n, _l.ptr(imgs_l), _l.ptr(imgs_r), w, h, imgs_l.strides[1],
            imgs_l.strides[0], mbf, mb, _l.ptr(frustums), local_map.ref(),
            None if map_index is None else _l.ptr(map_index), None if occupied is None else _l.ptr(occupied),
            C.byref(params), _l.ptr(out["kps_l"]), _l.ptr(out["desc_l"]), _l.ptr(out["n_l"]), _l.ptr(out["kps_r"]),
            _l.ptr(out["desc_r"]), _l.ptr(out["n_r"]), cap, _l.ptr(out["u_right"]), _l.ptr(out["depth"]),
            _l.ptr(out["n_matched"]), _l.ptr(out["assign"]), _l.ptr(out["nmatches"]), _l.ptr(out["n_in_view"])))
        return out

    @staticmethod
    def StereoTrackFramesBatchMulti(matchers, ex_lefts, ex_rights, imgs_l, imgs_r, mbf, mb, frustums=None, local_map=None,
                                    params=None, map_index=None, occupied=None, out=None):
        """orbm_stereo_track_frames_batch_multi: the pairs are sharded over len(matchers) handle sets (one per GPU) by
        the C library, one host thread per set. frustums = None gives the stereo-only form."""
        assert imgs_l.shape == imgs_r.shape and imgs_l.strides == imgs_r.strides and imgs_l.strides[2] == 1
        n, h, w = imgs_l.shape
        nd = len(matchers)
        cap = ex_lefts[0].capacity
        track = frustums is not None
        if out is None:
            out = (ORBmatcher.alloc_track_outputs if track else ORBmatcher.alloc_stereo_outputs)(n, cap)
        arr = lambda hs: (C.c_void_p * nd)(*[x._h for x in hs])
        L = matchers[0]._L
        rc = L.orbm_stereo_track_frames_batch_multi(
            nd, arr(matchers), arr(ex_lefts), arr(ex_rights), n, _l.ptr(imgs_l), _l.ptr(imgs_r), w, h, imgs_l.strides[1],
            imgs_l.strides[0], mbf, mb, _l.ptr(frustums) if track else None, local_map.ref() if track else None,
            None if map_index is None else _l.ptr(map_index), None if occupied is None else _l.ptr(occupied),
            C.byref(params) if track else None, _l.ptr(out["kps_l"]), _l.ptr(out["desc_l"]), _l.ptr(out["n_l"]),
            _l.ptr(out["kps_r"]), _l.ptr(out["desc_r"]), _l.ptr(out["n_r"]), cap, _l.ptr(out["u_right"]),
            _l.ptr(out["depth"]), _l.ptr(out["n_matched"]), _l.ptr(out["assign"]) if track else None,
            _l.ptr(out["nmatches"]) if track else None, _l.ptr(out["n_in_view"]) if track else None)
        if rc < 0:
            raise OrbxError(rc, "; ".join(L.orbm_last_error(m._h).decode() for m in matchers))
        return out

    # int SearchByProjection(Frame& F, const vector<MapPoint*>&, th, bFarPoints, thFarPoints) — src/ORBmatcher.cc:42
    def SearchByProjection(self, frame_view, mappoints, th=3.0, bFarPoints=False, thFarPoints=50.0):
        n = frame_view.struct.n
        assign = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_search_by_projection_map(self._h, frame_view.ref(), mappoints.ref(), th,
                                                          self.mfNNratio, int(bFarPoints), thFarPoints,
                                                          _l.ptr(assign), C.byref(nm)))
        return nm.value, assign[:n]

    # the same on a two-camera Frame (Nleft != -1): left search + right-camera twin + stereo partners — :42-221 incl. :148-217
    def SearchByProjectionFisheye(self, fisheye_view, mappoints, mappoints_right, th=3.0, bFarPoints=False,
                                  thFarPoints=50.0):
        n = fisheye_view.struct.n_left + fisheye_view.struct.n_right
        assign = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_search_by_projection_map_fisheye(self._h, fisheye_view.ref(), mappoints.ref(),
                                                                  mappoints_right.ref(), th, self.mfNNratio,
                                                                  int(bFarPoints), thFarPoints, _l.ptr(assign),
                                                                  C.byref(nm)))
        return nm.value, assign[:n]

    # void Frame::AssignFeaturesToGrid() — src/Frame.cc:520-547 (PosInGrid :833-844), on the device
    def AssignFeaturesToGrid(self, kps, min_x, min_y, inv_w, inv_h):
        kps = np.ascontiguousarray(kps)
        off = np.empty(64 * 48 + 1, np.int32)
        items = np.empty(max(len(kps), 1), np.int32)
        self._check(self._L.orbm_assign_features_to_grid(self._h, _l.ptr(kps), len(kps), min_x, min_y, inv_w, inv_h,
                                                         _l.ptr(off), _l.ptr(items)))
        return off, items[:off[-1]]

    # SearchByProjection(Frame&, vector<MapPoint*>) on frame `frame` of the extractor's last call, still on the device
    def SearchByProjectionResident(self, extractor, frame, n, mappoints, grid_params, u_right=None, occupied=None, th=3.0,
                                   bFarPoints=False, thFarPoints=50.0):
        assign = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        oc = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        min_x, min_y, inv_w, inv_h = grid_params
        self._check(self._L.orbm_search_by_projection_map_resident(
            self._h, extractor._h, frame, n, None if ur is None else _l.ptr(ur), None if oc is None else _l.ptr(oc),
            min_x, min_y, inv_w, inv_h, mappoints.ref(), th, self.mfNNratio, int(bFarPoints), thFarPoints,
            _l.ptr(assign), C.byref(nm)))
        return nm.value, assign[:n]

    # SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, ...) on a device-resident current frame
    def SearchByProjectionProjectedResident(self, extractor, frame, n, projected, grid_params, u_right=None,
                                            occupied=None, max_dist=TH_HIGH):
        assign = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        oc = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        min_x, min_y, inv_w, inv_h = grid_params
        self._check(self._L.orbm_search_by_projection_frame_resident(
            self._h, extractor._h, frame, n, None if ur is None else _l.ptr(ur), None if oc is None else _l.ptr(oc),
            min_x, min_y, inv_w, inv_h, projected.ref(), max_dist, int(self.mbCheckOrientation), _l.ptr(assign),
            C.byref(nm)))
        return nm.value, assign[:n]

    # int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) — :1594; KeyFrame form — :1808
    def SearchByProjectionProjected(self, frame_view, projected, max_dist=TH_HIGH):
        n = frame_view.struct.n
        assign = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_search_by_projection_frame(self._h, frame_view.ref(), projected.ref(), max_dist,
                                                            int(self.mbCheckOrientation), _l.ptr(assign),
                                                            C.byref(nm)))
        return nm.value, assign[:n]

    # one camera's candidate loop of the same function on a two-camera frame (:1649-1690 / :1711-1755): which keypoint
    # every point takes (orientation check left to the caller) and |GetFeaturesInArea| of its window
    def SearchByProjectionProjectedDecisions(self, frame_view, projected, max_dist=TH_HIGH):
        m = projected.struct.m
        dec, win = np.empty(max(m, 1), np.int32), np.empty(max(m, 1), np.int32)
        self._check(self._L.orbm_search_by_projection_frame_decisions(self._h, frame_view.ref(), projected.ref(),
                                                                      max_dist, _l.ptr(dec), _l.ptr(win)))
        return dec[:m], win[:m]

    # int SearchForTriangulation(KeyFrame*, KeyFrame*, vMatchedPairs, bOnlyStereo, bCoarse) — :886
    # Frame::ComputeBoW — src/Frame.cc:846-851: the vocabulary goes to the device once, then per-feature descents
    def SetVocabulary(self, voc):
        self._voc_keep = voc
        self._check(self._L.orbm_set_vocabulary(self._h, voc.ref()))

    def BowTransform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        w, wt, nd = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.float64), np.empty(max(n, 1), np.uint32)
        self._check(self._L.orbm_bow_transform(self._h, _l.ptr(desc), n, int(levelsup), _l.ptr(w), _l.ptr(wt),
                                               _l.ptr(nd)))
        return w[:n], wt[:n], nd[:n]

    # int SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize) — src/ORBmatcher.cc:618
    def SearchForInitialization(self, f1_view, f2_view, prev_matched_xy, windowSize=10):
        prev = np.ascontiguousarray(prev_matched_xy, np.float32).reshape(-1, 2)
        n = f1_view.struct.n
        m12 = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_search_for_initialization(self._h, f1_view.ref(), f2_view.ref(), _l.ptr(prev),
                                                           int(windowSize), self.mfNNratio,
                                                           int(self.mbCheckOrientation), _l.ptr(m12), C.byref(nm)))
        return nm.value, m12[:n]

    # the matching loop of int Fuse(KeyFrame* pKF, const vector<MapPoint*>&, th, bRight) — src/ORBmatcher.cc:1194-1257
    def FuseMatch(self, kf_view, inv_level_sigma2, projected, chi2_gate=True):
        inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
        m = projected.struct.m
        bi, bd = np.empty(max(m, 1), np.int32), np.empty(max(m, 1), np.int32)
        self._check(self._L.orbm_fuse_match(self._h, kf_view.ref(), _l.ptr(inv), projected.ref(), int(chi2_gate),
                                            _l.ptr(bi), _l.ptr(bd)))
        return bi[:m], bd[:m]

    # int SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) — src/ORBmatcher.cc:230
    def SearchByBoW(self, kf, frame):
        n = frame.struct.n
        mf = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_search_by_bow(self._h, kf.ref(), frame.ref(), self.mfNNratio,
                                               int(self.mbCheckOrientation), _l.ptr(mf), C.byref(nm)))
        return nm.value, mf[:n]

    # the same on a two-camera Frame (F.Nleft = n_left_frame != -1): left / right bests kept apart, :274-365
    def SearchByBoWTwoCameras(self, kf, frame, n_left_frame):
        n = frame.struct.n
        mf = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_search_by_bow_fisheye(self._h, kf.ref(), frame.ref(), int(n_left_frame), self.mfNNratio,
                                                       int(self.mbCheckOrientation), _l.ptr(mf), C.byref(nm)))
        return nm.value, mf[:n]

    # int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) — src/ORBmatcher.cc:766
    def SearchByBoWKeyFrames(self, kf1, kf2):
        n = kf1.struct.n
        m12 = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_search_by_bow_kf(self._h, kf1.ref(), kf2.ref(), self.mfNNratio,
                                                  int(self.mbCheckOrientation), _l.ptr(m12), C.byref(nm)))
        return nm.value, m12[:n]

    # the descriptor part of SearchForTriangulation for two-camera rigs (src/ORBmatcher.cc:973-988): CSR of candidates
    def TriangulationCandidates(self, kf1, kf2, cap=None):
        n1 = kf1.struct.n
        off = np.zeros(n1 + 1, np.int32)
        cap = int(cap if cap is not None else max(64 * n1, 1))
        idx, dist = np.empty(max(cap, 1), np.int32), np.empty(max(cap, 1), np.int32)
        total = C.c_int32(0)
        rc = self._L.orbm_triangulation_candidates(self._h, kf1.ref(), kf2.ref(), _l.ptr(off), _l.ptr(idx), _l.ptr(dist),
                                                   cap, C.byref(total))
        if rc == -2 and total.value > cap:   # ORBX_E_CAPACITY: retry with the reported size
            return self.TriangulationCandidates(kf1, kf2, total.value)
        self._check(rc)
        return off, idx[:total.value], dist[:total.value]

    def SearchForTriangulation(self, kf1, kf2, F12, ep, bOnlyStereo=False, bCoarse=False):
        F12 = np.ascontiguousarray(F12, np.float32).reshape(9)
        n = kf1.struct.n
        m12 = np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        self._check(self._L.orbm_search_for_triangulation(self._h, kf1.ref(), kf2.ref(), _l.ptr(F12), float(ep[0]),
                                                          float(ep[1]), int(bOnlyStereo), int(bCoarse),
                                                          int(self.mbCheckOrientation), _l.ptr(m12), C.byref(nm)))
        m12 = m12[:n]
        pairs = [(i, int(j)) for i, j in enumerate(m12) if j >= 0]  # vMatchedPairs, ascending idx1 (:1097-1103)
        return nm.value, m12, pairs
