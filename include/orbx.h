/* orbx.h — C ABI of the B200-native ORB front-end (liborbx.so).
 *
 * The reference (hellovuong/ORB_SLAM3_FAST @ 6255e16) has no plugin / FFI boundary: the hot path is the C++ class
 * surface of ORB_SLAM3::ORBextractor and ORB_SLAM3::ORBmatcher inside libORB_SLAM3.so. This header is the boundary a
 * maintainer binds instead: every entry point names the reference interface it replaces (paths relative to the
 * reference checkout). The header-compatible C++ classes that forward to it are in shim/ (see INTEGRATION.md).
 *
 * Conventions: plain pointers and sizes only; every call returns ORBX_OK (0) or a negative error code and never
 * throws or exits; a handle is used by one thread at a time (the reference runs its left and right extractor objects
 * on two threads, src/Frame.cc:200-203 — use two handles). All work of a handle is issued on its own CUDA stream.
 * There is NO CPU fallback: without a CUDA device every compute entry point returns ORBX_E_CUDA.
 */
#ifndef ORBX_H_
#define ORBX_H_

#include <stdint.h>

#include "orbx_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK 0
#define ORBX_E_EMPTY (-1)    /* empty image — ORBextractor::operator() returns -1 (src/ORBextractor.cc:1021) */
#define ORBX_E_CAPACITY (-2) /* an output buffer is too small; *n_out still holds the required count */
#define ORBX_E_ARG (-3)
#define ORBX_E_CUDA (-4)     /* CUDA error or no device; see orbx_last_error */
#define ORBX_E_SIZE (-5)     /* image too small for the level count (the reference divides by zero,
                                src/ORBextractor.cc:781-784) or larger than 4096 px */

typedef struct orbx_extractor orbx_extractor;

/* ---- ORBextractor (include/ORBextractor.h:48-120) ---------------------------------------------------------- */

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) (include/ORBextractor.h:53-57,
 * src/ORBextractor.cc:408-469). `device` = CUDA ordinal; `max_batch` = frames per launch group of the batched calls
 * (device scratch is sized for it lazily, when the first image size is seen). */
int orbx_extractor_create(orbx_extractor** out, int device, int nfeatures, float scale_factor, int nlevels,
                          int ini_th_fast, int min_th_fast, int max_batch);
void orbx_extractor_destroy(orbx_extractor* ex);
/* Text of the last error on this handle (or of the last failed create when ex == NULL). */
const char* orbx_last_error(const orbx_extractor* ex);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:70-84) and mnFeaturesPerLevel (:114). Arrays of nlevels entries; any pointer may be NULL. */
int orbx_extractor_levels(const orbx_extractor* ex);
int orbx_extractor_tables(const orbx_extractor* ex, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int32_t* features_per_level);
/* Rows an output buffer needs per frame: the quadtree may return up to 3 keypoints more than each level's quota
 * (src/ORBextractor.cc:729), so N can exceed nfeatures. */
int orbx_extractor_capacity(const orbx_extractor* ex);

/* int ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea) (include/ORBextractor.h:64-68,
 * src/ORBextractor.cc:1015-1106). image: 8-bit single channel, `stride` bytes per row, host memory. kps / desc:
 * caller-allocated, `cap` rows (cv::KeyPoint layout / 32-byte descriptor rows). *n_out = number of keypoints,
 * *mono_index = the reference's return value. Returns ORBX_E_EMPTY for an empty image like the reference's -1. */
int orbx_extract(orbx_extractor* ex, const uint8_t* image, int width, int height, int stride, int lap0, int lap1,
                 orbx_kp* kps, uint8_t* desc, int cap, int32_t* n_out, int32_t* mono_index);

/* Batched operator(): n_frames images of one size, `frame_stride` bytes apart, host memory (pinned memory makes the
 * copies asynchronous: orbx_host_alloc). Outputs are [n_frames][cap] rows, n_out / mono_index [n_frames].
 * Frames are independent (the extractor keeps no state between frames). Returns the first error. */
int orbx_extract_batch(orbx_extractor* ex, int n_frames, const uint8_t* images, int width, int height, int stride,
                       int64_t frame_stride, int lap0, int lap1, orbx_kp* kps, uint8_t* desc, int cap,
                       int32_t* n_out, int32_t* mono_index);

/* orbx_extract_batch over several GPUs of one box from ONE C/C++ caller ("frames sharded over
 * visible GPUs"): ex[d] is an extractor created on device d' (any ordinals; same parameters), n_devices >= 1. Frames are
 * independent, so they are cut into n_devices contiguous blocks (sizes differ by at most one, the first blocks take the
 * extra) and every block runs orbx_extract_batch on its own host thread with its own handle; there is no inter-GPU
 * traffic. Outputs land at the frames' global positions. Returns the first error. */
int orbx_extract_batch_multi(int n_devices, orbx_extractor* const* ex, int n_frames, const uint8_t* images, int width,
                             int height, int stride, int64_t frame_stride, int lap0, int lap1, orbx_kp* kps,
                             uint8_t* desc, int cap, int32_t* n_out, int32_t* mono_index);

/* Same, inputs and outputs resident in device memory; enqueued on `cuda_stream` (a cudaStream_t; NULL = the handle's
 * stream) and NOT synchronised. n_frames <= max_batch. d_status[n_frames] receives 0 or ORBX_E_CAPACITY per frame.
 * Stream contract: the call works in the handle's scratch (pyramid, blurred levels, candidate lists of lane 0), which
 * the host-facing calls use on the handle's own stream. A handle therefore serves ONE stream at a time: two device
 * calls on different streams, or a device call followed by a host-facing call (or by anything that reads the frame
 * "of the last call": orbx_download_pyramid, orbm_stereo_match_batch_device, the *_resident searches), must be ordered
 * by the caller — an event, or a synchronise — exactly like two kernels that share a buffer. The library adds no event
 * of its own here, so that the call can sit inside a stream capture of the caller. */
int orbx_extract_batch_device(orbx_extractor* ex, int n_frames, const uint8_t* d_images, int width, int height,
                              int stride, int64_t frame_stride, int lap0, int lap1, orbx_kp* d_kps, uint8_t* d_desc,
                              int cap, int32_t* d_n, int32_t* d_mono_index, int32_t* d_status, void* cuda_stream);

/* std::vector<cv::Mat> mvImagePyramid (include/ORBextractor.h:86), read by Frame::ComputeStereoMatches
 * (src/Frame.cc:927,1011,1024,1029): level size, and a host copy of one level of frame `frame` of the last call in
 * the reference's layout — the (w + 38) x (h + 38) buffer with the 19-px BORDER_REFLECT_101 frame
 * (src/ORBextractor.cc:1114-1143); the cv::Mat the shim exposes is the ROI at (19, 19).
 * `frame` (here, in the orbx_debug_* calls and in the orbm_* calls that take an extractor + frame) is the index of the
 * frame inside the extractor's most recent extract call. A call keeps the device state of its last 8 groups of
 * max_batch frames: every frame of a call of <= 8 * max_batch frames is addressable, of a longer call only the tail;
 * a frame that is gone (or an index from an earlier call) gives ORBX_E_ARG, never another frame's data. */
int orbx_level_size(const orbx_extractor* ex, int level, int* width, int* height);
int orbx_download_pyramid(orbx_extractor* ex, int frame, int level, uint8_t* dst, int dst_stride);

/* Stage outputs of the last call, for stage-by-stage parity tests: which = 0 raw level, 1 blurred level (w x h). */
int orbx_debug_level(orbx_extractor* ex, int frame, int level, int which, uint8_t* dst, int dst_stride);
/* FAST candidates before the quadtree (x, y relative to minBorder as in vToDistributeKeys) / keypoints after it
 * (level coordinates). Returns the count (or a negative error); fills at most cap entries. */
int orbx_debug_candidates(orbx_extractor* ex, int frame, int level, orbx_kp* out, int cap);
int orbx_debug_level_keypoints(orbx_extractor* ex, int frame, int level, orbx_kp* out, int cap);

/* Per-stage device time of the calls since the last reset, measured with CUDA events on the handle's stream when
 * profiling is enabled (adds event records between kernels; leave it off for throughput runs).
 * Stages: 0 pyramid, 1 fast, 2 quadtree, 3 blur, 4 describe. ms[5] accumulates, launches[5] counts. */
int orbx_profile_enable(orbx_extractor* ex, int on);
int orbx_profile_read(orbx_extractor* ex, float* ms, int32_t* launches, int reset);

/* cv::remap(src, dst, M1, M2, cv::INTER_LINEAR) for 8-bit single-channel images and CV_32FC1 maps (border constant 0):
 * the stereo rectification System::TrackStereo applies before the Frame constructor (src/System.cc:286-294; the maps
 * are built once by cv::initUndistortRectifyMap(..., CV_32F, ...), src/Settings.cc:557-572) — SURVEY.md §8(f) rank 4.
 * mapx / mapy: [dst_height][dst_width] floats, dense. The device form rectifies n_frames images with the same maps on
 * `cuda_stream` without synchronising (its output can be handed to orbx_extract_batch_device); the host form is
 * synchronous. */
int orbx_remap_linear_device(int device, int n_frames, const uint8_t* d_src, int src_width, int src_height,
                             int src_stride, int64_t src_frame_stride, const float* d_mapx, const float* d_mapy,
                             int dst_width, int dst_height, uint8_t* d_dst, int dst_stride, int64_t dst_frame_stride,
                             void* cuda_stream);
int orbx_remap_linear(int device, const uint8_t* src, int src_width, int src_height, int src_stride, const float* mapx,
                      const float* mapy, int dst_width, int dst_height, uint8_t* dst, int dst_stride);

/* cv::cvtColor(src, dst, cv::COLOR_{BGR,RGB,BGRA,RGBA}2GRAY) on 8-bit images — what Tracking::GrabImageStereo /
 * GrabImageRGBD / GrabImageMonocular apply to colour input before the Frame constructor (src/Tracking.cc:1394-1412,
 * 1500-1513, 1558-1571); SURVEY.md §8(f) rank 4 (image front-end). channels = 3 or 4, rgb != 0 when the first channel
 * is red (mbRGB). The device form converts n_frames images in place of the caller's buffers on `cuda_stream` (NULL =
 * the default stream) without synchronising, so a colour batch can feed orbx_extract_batch_device directly; the host
 * form is synchronous. No handle: errors are return codes only. */
int orbx_cvt_gray_device(int device, int n_frames, const uint8_t* d_src, int width, int height, int src_stride,
                         int64_t src_frame_stride, int channels, int rgb, uint8_t* d_dst, int dst_stride,
                         int64_t dst_frame_stride, void* cuda_stream);
int orbx_cvt_gray(int device, const uint8_t* src, int width, int height, int src_stride, int channels, int rgb,
                  uint8_t* dst, int dst_stride);

/* Number of kernels one extract call launches for the geometry the handle is planned for (0 before the first call):
 * one resize per level >= 1, one or two FAST launches (the small levels' tall cells get their own shared-memory
 * layout), quadtree, blur, describe. bench.py reports it as gpu_launches. */
int orbx_kernel_launches(const orbx_extractor* ex);

/* Pinned host memory for the batched calls. */
void* orbx_host_alloc(int64_t bytes);
void orbx_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H_ */
