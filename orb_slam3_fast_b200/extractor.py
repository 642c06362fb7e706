"""Host-side mirror of ORB_SLAM3::ORBextractor (include/ORBextractor.h:48-120) over the orbx C ABI.

Same constructor arguments, same call semantics (`operator()` returns monoIndex and fills keypoints + descriptors,
-1 on an empty image), same getters, plus the batched calls the B200 path adds. numpy arrays stand in for
std::vector<cv::KeyPoint> (a structured dtype with cv::KeyPoint's 28-byte layout) and the N x 32 CV_8U cv::Mat.
"""
import ctypes as C

import numpy as np

from . import lib as _l
from .lib import KP_DTYPE, OrbxError

STAGES = ("pyramid", "fast", "quadtree", "blur", "describe")


class ORBextractor:
    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, device=0, max_batch=1):
        self._L = _l.lib()
        h = C.c_void_p()
        rc = self._L.orbx_extractor_create(C.byref(h), device, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                                           max_batch)
        if rc != 0:
            raise OrbxError(rc, self._L.orbx_last_error(None).decode())
        self._h = h
        self.nfeatures, self.nlevels, self.max_batch, self.device = nfeatures, nlevels, max_batch, device
        self._scale = np.empty(nlevels, np.float32)
        self._inv_scale = np.empty(nlevels, np.float32)
        self._sigma2 = np.empty(nlevels, np.float32)
        self._inv_sigma2 = np.empty(nlevels, np.float32)
        self.mnFeaturesPerLevel = np.empty(nlevels, np.int32)
        self._check(self._L.orbx_extractor_tables(h, _l.ptr(self._scale), _l.ptr(self._inv_scale), _l.ptr(self._sigma2),
                                                  _l.ptr(self._inv_sigma2), _l.ptr(self.mnFeaturesPerLevel)))
        self.capacity = self._L.orbx_extractor_capacity(h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbx_extractor_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise OrbxError(rc, self._L.orbx_last_error(self._h).decode())
        return rc

    # ---- getters (include/ORBextractor.h:70-84) ----
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return float(self._scale[1]) if self.nlevels > 1 else 1.0

    def GetScaleFactors(self):
        return self._scale

    def GetInverseScaleFactors(self):
        return self._inv_scale

    def GetScaleSigmaSquares(self):
        return self._sigma2

    def GetInverseScaleSigmaSquares(self):
        return self._inv_sigma2

    # ---- operator() ----
    def __call__(self, image, vLappingArea=(0, 0)):
        """Returns (monoIndex, keypoints, descriptors); monoIndex == -1 and empty outputs for an empty image."""
        if image is None or image.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        cap = self.capacity
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n, mono = C.c_int32(0), C.c_int32(0)
        self._check(self._L.orbx_extract(self._h, _l.ptr(image), image.shape[1], image.shape[0], image.strides[0],
                                         int(vLappingArea[0]), int(vLappingArea[1]), _l.ptr(kps), _l.ptr(desc), cap,
                                         C.byref(n), C.byref(mono)))
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, vLappingArea=(0, 0), out=None):
        """images: uint8 [n, h, w] host array. Returns (n_out[n], mono_index[n], kps[n, cap], desc[n, cap, 32])."""
        images = np.ascontiguousarray(images, np.uint8)
        nf, h, w = images.shape
        cap = self.capacity
        if out is None:
            out = (np.empty(nf, np.int32), np.empty(nf, np.int32), np.empty((nf, cap), KP_DTYPE),
                   np.empty((nf, cap, 32), np.uint8))
        n_out, mono, kps, desc = out
        self._check(self._L.orbx_extract_batch(self._h, nf, _l.ptr(images), w, h, images.strides[1], images.strides[0],
                                               int(vLappingArea[0]), int(vLappingArea[1]), _l.ptr(kps), _l.ptr(desc),
                                               cap, _l.ptr(n_out), _l.ptr(mono)))
        return n_out, mono, kps, desc

    @staticmethod
    def extract_batch_multi(extractors, images, vLappingArea=(0, 0), out=None):
        """orbx_extract_batch_multi: the frames are sharded over the given extractors (one per GPU) by the C library, one
        host thread each. Returns (n_out[n], mono_index[n], kps[n, cap], desc[n, cap, 32])."""
        images = np.ascontiguousarray(images, np.uint8)
        nf, h, w = images.shape
        cap = extractors[0].capacity
        if out is None:
            out = (np.empty(nf, np.int32), np.empty(nf, np.int32), np.empty((nf, cap), KP_DTYPE),
                   np.empty((nf, cap, 32), np.uint8))
        n_out, mono, kps, desc = out
        L = extractors[0]._L
        L.orbx_extract_batch_multi.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                               C.c_void_p]
        hs = (C.c_void_p * len(extractors))(*[e._h for e in extractors])
        rc = L.orbx_extract_batch_multi(len(extractors), hs, nf, _l.ptr(images), w, h, images.strides[1], images.strides[0],
                                        int(vLappingArea[0]), int(vLappingArea[1]), _l.ptr(kps), _l.ptr(desc), cap,
                                        _l.ptr(n_out), _l.ptr(mono))
        if rc < 0:
            raise OrbxError(rc, "; ".join(L.orbx_last_error(e._h).decode() for e in extractors))
        return n_out, mono, kps, desc

    def extract_batch_device(self, d_images, n_frames, w, h, stride, frame_stride, vLappingArea, d_kps, d_desc, cap,
                             d_n, d_mono, d_status, stream=0):
        """All pointers are device addresses (ints). Enqueues on `stream` (a cudaStream_t value, 0 = handle stream)."""
        self._check(self._L.orbx_extract_batch_device(self._h, n_frames, C.c_void_p(d_images), w, h, stride,
                                                      frame_stride, int(vLappingArea[0]), int(vLappingArea[1]),
                                                      C.c_void_p(d_kps), C.c_void_p(d_desc), cap, C.c_void_p(d_n),
                                                      C.c_void_p(d_mono), C.c_void_p(d_status),
                                                      C.c_void_p(stream) if stream else None))

    # ---- mvImagePyramid (include/ORBextractor.h:86) ----
    def level_size(self, level):
        w, h = C.c_int(0), C.c_int(0)
        self._check(self._L.orbx_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def image_pyramid_bordered(self, level, frame=0):
        w, h = self.level_size(level)
        buf = np.empty((h + 38, w + 38), np.uint8)
        self._check(self._L.orbx_download_pyramid(self._h, frame, level, _l.ptr(buf), buf.strides[0]))
        return buf

    def image_pyramid(self, level, frame=0):
        """The cv::Mat the reference keeps in mvImagePyramid[level]: the ROI at (19, 19) of the bordered buffer."""
        return self.image_pyramid_bordered(level, frame)[19:-19, 19:-19]

    # ---- stage outputs for parity tests ----
    def debug_level(self, level, blurred=False, frame=0):
        w, h = self.level_size(level)
        buf = np.empty((h, w), np.uint8)
        self._check(self._L.orbx_debug_level(self._h, frame, level, 1 if blurred else 0, _l.ptr(buf), buf.strides[0]))
        return buf

    def debug_candidates(self, level, frame=0):
        n = self._check(self._L.orbx_debug_candidates(self._h, frame, level, None, 0))
        out = np.empty(max(n, 1), KP_DTYPE)
        self._check(self._L.orbx_debug_candidates(self._h, frame, level, _l.ptr(out), n))
        return out[:n]

    def debug_level_keypoints(self, level, frame=0):
        n = self._check(self._L.orbx_debug_level_keypoints(self._h, frame, level, None, 0))
        out = np.empty(max(n, 1), KP_DTYPE)
        self._check(self._L.orbx_debug_level_keypoints(self._h, frame, level, _l.ptr(out), n))
        return out[:n]

    # ---- per-stage device timing ----
    def profile(self, on=True):
        self._check(self._L.orbx_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset=True):
        ms = np.zeros(len(STAGES), np.float32)
        cnt = np.zeros(len(STAGES), np.int32)
        self._check(self._L.orbx_profile_read(self._h, _l.ptr(ms), _l.ptr(cnt), 1 if reset else 0))
        return dict(zip(STAGES, ms.tolist())), dict(zip(STAGES, cnt.tolist()))


def cvtColorToGray(img, rgb=False, device=0):
    """cv::cvtColor(img, COLOR_{BGR,RGB,BGRA,RGBA}2GRAY) (src/Tracking.cc:1394-1412) for an [h, w, 3|4] uint8 host image."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, c = img.shape
    out = np.empty((h, w), np.uint8)
    L = _l.lib()
    rc = L.orbx_cvt_gray(device, _l.ptr(img), w, h, img.strides[0], c, int(rgb), _l.ptr(out), w)
    if rc != 0:
        raise OrbxError(rc, "orbx_cvt_gray failed")
    return out


def remapLinear(img, mapx, mapy, device=0):
    """cv::remap(img, mapx, mapy, INTER_LINEAR) (src/System.cc:293-294) for a uint8 [h, w] host image, float32 maps."""
    img = np.ascontiguousarray(img, np.uint8)
    mapx, mapy = np.ascontiguousarray(mapx, np.float32), np.ascontiguousarray(mapy, np.float32)
    dh, dw = mapx.shape
    out = np.empty((dh, dw), np.uint8)
    rc = _l.lib().orbx_remap_linear(device, _l.ptr(img), img.shape[1], img.shape[0], img.strides[0], _l.ptr(mapx),
                                    _l.ptr(mapy), dw, dh, _l.ptr(out), dw)
    if rc != 0:
        raise OrbxError(rc, "orbx_remap_linear failed")
    return out
