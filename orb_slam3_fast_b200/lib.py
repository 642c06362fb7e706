"""ctypes loader for liborbx.so (the C ABI of include/orbx.h and include/orbm.h).

There is no Python or CPU implementation behind this module: if the shared library is missing the import fails with
instructions to build it, and every compute entry point returns ORBX_E_CUDA when no CUDA device is present.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ORBX_SO_PATH") or os.path.join(_HERE, "liborbx.so")  # the override is a development aid
# (instrumented builds, e.g. -DORBX_QT_PROF); there is still no fallback: a missing file fails the import
CSRC = os.path.join(_HERE, "csrc")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28  # cv::KeyPoint

OK, E_EMPTY, E_CAPACITY, E_ARG, E_CUDA, E_SIZE = 0, -1, -2, -3, -4, -5


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("orbx error %d: %s" % (code, msg))
        self.code = code


def build(force=False):
    """Compiles liborbx.so in-tree with nvcc for sm_100a (no GPU needed)."""
    args = ["make", "-C", CSRC, "-s", "-j8"] + (["-B"] if force else [])
    subprocess.check_call(args)
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C %s). "
                          "orb_slam3_fast_b200 has no CPU fallback." % (SO_PATH, CSRC))
    L = C.CDLL(SO_PATH)
    vp, ci, cf, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    L.orbx_extractor_create.argtypes = [C.POINTER(vp), ci, ci, cf, ci, ci, ci, ci]
    L.orbx_extractor_destroy.argtypes = [vp]
    L.orbx_extractor_destroy.restype = None
    L.orbx_last_error.argtypes = [vp]
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_extractor_levels.argtypes = [vp]
    L.orbx_extractor_tables.argtypes = [vp] * 6
    L.orbx_extractor_capacity.argtypes = [vp]
    L.orbx_extract.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp]
    L.orbx_extract_batch.argtypes = [vp, ci, vp, ci, ci, ci, i64, ci, ci, vp, vp, ci, vp, vp]
    L.orbx_extract_batch_device.argtypes = [vp, ci, vp, ci, ci, ci, i64, ci, ci, vp, vp, ci, vp, vp, vp, vp]
    L.orbx_level_size.argtypes = [vp, ci, vp, vp]
    L.orbx_download_pyramid.argtypes = [vp, ci, ci, vp, ci]
    L.orbx_debug_level.argtypes = [vp, ci, ci, ci, vp, ci]
    L.orbx_debug_candidates.argtypes = [vp, ci, ci, vp, ci]
    L.orbx_debug_level_keypoints.argtypes = [vp, ci, ci, vp, ci]
    L.orbx_profile_enable.argtypes = [vp, ci]
    L.orbx_profile_read.argtypes = [vp, vp, vp, ci]
    L.orbx_remap_linear.argtypes = [ci, vp, ci, ci, ci, vp, vp, ci, ci, vp, ci]
    L.orbx_remap_linear_device.argtypes = [ci, ci, vp, ci, ci, ci, i64, vp, vp, ci, ci, vp, ci, i64, vp]
    L.orbx_cvt_gray.argtypes = [ci, vp, ci, ci, ci, ci, ci, vp, ci]
    L.orbx_cvt_gray_device.argtypes = [ci, ci, vp, ci, ci, ci, i64, ci, ci, vp, ci, i64, vp]
    L.orbx_kernel_launches.argtypes = [vp]
    L.orbx_host_alloc.argtypes = [i64]
    L.orbx_host_alloc.restype = vp
    L.orbx_host_free.argtypes = [vp]
    L.orbx_host_free.restype = None
    _lib = L
    return L


def ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)
