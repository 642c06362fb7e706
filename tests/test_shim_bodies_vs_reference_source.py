"""The reference-side drop-in bodies (shim/ORBmatcher_orbx.cc, shim/ORBmatcher_next_orbx.cc) against the reference's own
ORBmatcher.cc, on the CPU: oracle/_ref/libshim_world.so links the shim bodies IN PLACE OF the reference's methods of the
same name (they are listed first; the linker keeps the first definition), over the stand-in Frame / KeyFrame / MapPoint
world, with the orbm C ABI answered by the oracle (tests/mock_orbm_oracle.cpp) instead of the CUDA library. What this
exercises is the shim's glue: flattening the pointer graph into the views, the host-side projection, scattering the
answers back. The same comparisons as tests/test_oracle_matchers_vs_reference_source.py are then run through it."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import refsrc
import test_oracle_matchers_vs_reference_source as T

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libshim_world.so")
pytestmark = pytest.mark.skipif(not os.path.exists(_PATH), reason="oracle/_ref/libshim_world.so not built")


@pytest.fixture()
def shim_world(monkeypatch):
    lib = C.CDLL(_PATH)
    for name in ("orbrefsrc_search_by_projection_map", "orbrefsrc_search_for_triangulation", "orbrefsrc_search_by_bow",
                 "orbrefsrc_search_by_bow_kf", "orbrefsrc_search_by_projection_last_frame", "orbrefsrc_fuse",
                 "orbrefsrc_fuse_two_camera", "orbrefsrc_search_by_projection_last_frame_fisheye",
                 "orbrefsrc_features_in_area", "orbrefsrc_stereo_frame", "orbrefsrc_search_for_initialization",
                 "orbrefsrc_search_by_projection_keyframe", "orbrefsrc_search_by_projection_sim3", "orbrefsrc_search_by_sim3",
                 "orbrefsrc_distinctive_descriptor", "orbrefsrc_search_by_projection_map_fisheye",
                 "orbrefsrc_search_by_bow_fisheye", "orbrefsrc_stereo_fisheye",
                 "orbrefsrc_search_by_bow_kf_fisheye", "orbrefsrc_search_for_triangulation_fisheye"):
        getattr(lib, name).restype = C.c_int
    refsrc.mlib()
    monkeypatch.setattr(refsrc, "_mlib", lib)
    return lib


@pytest.mark.parametrize("args", [(True, 1.0, 0.8, True, 3), (False, 3.0, 0.8, False, 4), (True, 5.0, 0.6, True, 5)])
def test_shim_search_by_projection_map(shim_world, args):
    T.test_search_by_projection_map(*args)


@pytest.mark.parametrize("args", [(1, True, 7.0, True, 0), (-1, True, 7.0, True, 1), (0, True, 15.0, True, 2),
                                  (0, False, 15.0, False, 3)])
def test_shim_search_by_projection_last_frame(shim_world, args):
    T.test_search_by_projection_last_frame(*args)


@pytest.mark.parametrize("args", [(False, False, True, (1e6, 200.0)), (True, False, True, (1e6, 200.0)),
                                  (False, True, False, (320.0, 200.0)), (False, False, False, (-5000.0, 200.0))])
def test_shim_search_for_triangulation(shim_world, args):
    # the shim computes F12 itself (K1^-T [t12]x R12 K2^-1, src/CameraModels/Pinhole.cpp:130-133) from the KeyFrame poses:
    # identity intrinsics and rotation in the stand-in world, t12 = (-ep_x, -ep_y, 0)  =>  F12 = [t12]x
    ex, ey = args[3]
    F12 = np.array([[0, 0, -ey], [0, 0, ex], [ey, -ex, 0]], np.float32)
    T.test_search_for_triangulation(*args, F12=F12)


@pytest.mark.parametrize("args", [(1, 0.7, True), (2, 0.9, False), (3, 0.6, True)])
def test_shim_search_by_bow(shim_world, args):
    T.test_search_by_bow_both_overloads(*args)


@pytest.mark.parametrize("args", [(1, 0.7, True, True), (2, 0.9, False, False), (3, 0.6, True, True)])
def test_shim_search_by_bow_two_camera_frame(shim_world, args):
    T.test_search_by_bow_two_camera_frame(*args)


@pytest.mark.parametrize("args", [(False, False, True, 0.5, 0.6), (False, False, False, 0.3, 1.0), (False, True, True, 0.5, 0.5),
                                  (True, False, True, 0.5, 0.5), (False, False, True, 0.0, 0.4)])
def test_shim_search_for_triangulation_two_camera_keyframes(shim_world, args):
    T.test_search_for_triangulation_two_camera_keyframes(*args)


@pytest.mark.parametrize("args", [(1, 0.7, True), (2, 0.9, False)])
def test_shim_search_by_bow_two_camera_keyframes(shim_world, args):
    T.test_search_by_bow_two_camera_keyframes(*args)


@pytest.mark.parametrize("args", [(0, 7.0, True, 20), (1, 7.0, True, 21), (-1, 10.0, True, 22), (0, 15.0, False, 23)])
def test_shim_search_by_projection_last_frame_two_camera(shim_world, args):
    T.test_search_by_projection_last_frame_two_camera(*args)


@pytest.mark.parametrize("args", [(False, 3.0, 14), (True, 3.0, 15), (True, 2.5, 16)])
def test_shim_fuse_two_camera_keyframe(shim_world, args):
    T.test_fuse_two_camera_keyframe(*args)


@pytest.mark.parametrize("args", [(1, 0, 0), (2, 150, 90), (3, 399, 0), (4, 0, 398)])
def test_shim_compute_stereo_fisheye_matches(shim_world, args):
    T.test_compute_stereo_fisheye_matches(*args)


@pytest.mark.parametrize("args", [(False, 3.0, 4), (False, 2.5, 6)])
def test_shim_fuse(shim_world, args):
    T.test_fuse_both_overloads(*args)


def test_shim_assign_features_to_grid(shim_world):
    T.test_grid_functions()


@pytest.mark.parametrize("args", [(752, 480, 1200, 1, "scene"), (400, 300, 600, 8, "scene")])
def test_shim_extractor_class_and_compute_stereo_matches(shim_world, args):
    """shim/ORBextractor.{h,cc} (constructor tables, operator(), the mvImagePyramid mirror) and the drop-in
    Frame::ComputeStereoMatches of shim/FrameStereo_orbx.cc, driven exactly like the reference's stereo Frame constructor
    drives the originals: keypoints, descriptors, mvuRight and mvDepth must come back as the oracle's."""
    T.test_stereo_frame_hot_path_equals_the_reference_source(*args)


# ---- shim/ORBmatcher_sim3_orbx.cc: the Sim3 family, SearchForInitialization, the KeyFrame-set projection, distinctive descriptors ----
@pytest.mark.parametrize("args", [(False, 8, 1.0, 8), (True, 8, 1.0, 9), (False, 4, 1.5, 10)])
def test_shim_sim3_search_by_projection(shim_world, args):
    T.test_sim3_search_by_projection_is_the_projected_form(*args)


@pytest.mark.parametrize("args", [(7.5, 11), (4.0, 12)])
def test_shim_search_by_sim3(shim_world, args):
    T.test_search_by_sim3_is_two_gate_free_fuse_matches_plus_agreement(*args)


@pytest.mark.parametrize("args", [(True, 3.0, 5), (True, 4.0, 7)])
def test_shim_fuse_sim3(shim_world, args):
    T.test_fuse_both_overloads(*args)


@pytest.mark.parametrize("args", [(6, 30, 0.9, True), (7, 60, 0.9, False), (8, 100, 0.7, True)])
def test_shim_search_for_initialization(shim_world, args):
    T.test_search_for_initialization(*args)


@pytest.mark.parametrize("args", [(10.0, 100, True, 4), (3.0, 64, True, 5), (10.0, 100, False, 6)])
def test_shim_search_by_projection_keyframe(shim_world, args):
    T.test_search_by_projection_keyframe(*args)


def test_shim_compute_distinctive_descriptors(shim_world):
    T.test_compute_distinctive_descriptors()


@pytest.mark.parametrize("args", [(0, 1.0, True, (700, 650, 3000)), (2, 6.0, True, (1200, 1100, 10000))])
def test_shim_search_by_projection_two_cameras(shim_world, args):
    """the fisheye (Nleft != -1) path of the drop-in SearchByProjection(Frame&, vector<MapPoint*>)"""
    T.test_search_by_projection_map_fisheye(*args)


# ---- shim/Tracking_orbx.cc: the caller of the local-map search ----
@pytest.mark.parametrize("args", [(0, 10000, 0), (1, 4000, 1), (2, 4000, 2), (3, 3000, 3), (4, 3000, 4), (5, 3000, 5),
                                  (6, 2000, 6), (7, 2000, 7), (8, 1, 0), (9, 0, 0)])
def test_shim_tracking_search_local_points(shim_world, args):
    """the drop-in Tracking::SearchLocalPoints (one orbm_is_in_frustum call for the whole local map + the drop-in
    SearchByProjection): mvpMapPoints, every MapPoint tracking word, mnVisible, mnLastFrameSeen and mmProjectPoints as the
    reference's own function text leaves them"""
    T.test_search_local_points(*args)


def test_shim_bodies_from_three_threads(shim_world):
    """ORBmatcher temporaries live on the Tracking, LocalMapping and LoopClosing threads at once (src/Tracking.cc:3303,
    src/LocalMapping.cc:435, src/LoopClosing.cc:729): three threads run drop-in bodies concurrently (ctypes releases the
    GIL during the calls) — SearchLocalPoints + SearchByProjection on different frames and maps of one image size (the
    stand-in Frame keeps the image bounds in statics, like the reference). Each thread's results must be the serial ones:
    the matcher context is per thread (shim/orbx_thread_matcher.h) and the bodies keep no other state."""
    import threading
    errors = []
    T.test_search_local_points(8, 1, 0)   # one serial call first: the stand-in Frame's static image bounds get their values

    def worker(t):
        try:
            for seed, m, case in ((t, 3000, t % 3), (t + 3, 2000, 5 + t % 3)):
                T.test_search_local_points(seed, m, case)
            T.test_search_by_projection_map(True, 1.0 + 2 * t, 0.8, bool(t & 1), 3 + t)
        except BaseException as e:  # noqa: BLE001 - an assertion of the comparison, or whatever the thread died of
            errors.append((t, repr(e)[:300]))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(3)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors



def test_shim_tracking_search_local_points_two_camera_frame(shim_world):
    """the Nleft != -1 branch of the drop-in body (per-point Frame::isInFrustum kept on the host)"""
    T.test_search_local_points_two_camera_frame()
