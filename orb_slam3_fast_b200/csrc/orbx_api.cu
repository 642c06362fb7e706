// orbx_api.cu — the extern "C" extractor ABI declared in include/orbx.h: handle, device memory, streams, launch order.
// No result is ever computed on the host: a missing device or a CUDA error is reported as ORBX_E_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/orbx.h"
#include "orbx_handle.h"
#include "orbx_kernels.cuh"
#include "orbx_quadtree.h"

using namespace orbx;

// The pipelined host-facing calls keep 2 streams per lane busy (8 in all) next to the caller's own; with the default of
// 8 hardware work queues streams alias and falsely serialise (measured: 109 k -> 115 k frames/s end to end with 32). The
// variable is only read when the CUDA context is created, so it is set — if the process has not chosen a value — when
// the library is loaded.
__attribute__((constructor)) static void orbx_default_connections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

namespace {
const int8_t kPatternHost[256 * 4] = {
#include "orb_pattern.inc"
};
thread_local std::string g_create_error;

#define ORBX_CUDA(ex, call)                                                                              \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return api_fail(ex, ORBX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

void free_plan_buffers(orbx_extractor* ex) {
  orbx::alloc_generation()++;
  for (OrbxLane& L : ex->lane) {
    cudaFree(L.d_in);
    cudaFree(L.d_pyr);
    cudaFree(L.d_blur);
    cudaFree(L.ws.slots);
    cudaFree(L.ws.cell_count);
    cudaFree(L.ws.cand);
    cudaFree(L.ws.lab);
    cudaFree(L.ws.lvl_kp);
    cudaFree(L.ws.lvl_n);
    cudaFree(L.ws.lvl_c);
    cudaFree(L.ws.lvl_st);
    cudaFree(L.ws.qt_prof);
    cudaFree(L.ws.dst);
    L.d_in = L.d_pyr = L.d_blur = nullptr;
    L.ws = WorkSet{};
    L.last_frames = 0;
  }
  cudaFree(ex->d_tab);
  ex->d_tab = nullptr;
  ex->planned = false;
}

void free_out_buffers(orbx_extractor* ex) {
  orbx::alloc_generation()++;
  for (OrbxLane& L : ex->lane) {
    cudaFree(L.d_kps);
    cudaFree(L.d_desc);
    cudaFree(L.d_n);
    cudaFree(L.d_mono);
    cudaFree(L.d_status);
    L.d_kps = nullptr;
    L.d_desc = nullptr;
    L.d_n = L.d_mono = L.d_status = nullptr;
    L.out_cap = 0;
  }
}

int sync_all_lanes(orbx_extractor* ex) {
  for (OrbxLane& L : ex->lane)
    if (L.stream) ORBX_CUDA(ex, cudaStreamSynchronize(L.stream));
  return ORBX_OK;
}

// scratch of one lane; lanes beyond the first are only materialised when a call needs them
int alloc_lane(orbx_extractor* ex, OrbxLane& L) {
  if (L.d_pyr) return ORBX_OK;
  orbx::alloc_generation()++;
  const Plan& P = ex->plan;
  const size_t B = ex->max_batch;
  ORBX_CUDA(ex, cudaMalloc(&L.d_in, (size_t)ex->in_fstride * B));
  ORBX_CUDA(ex, cudaMalloc(&L.d_pyr, (size_t)ex->slab_fstride * B));
  ORBX_CUDA(ex, cudaMalloc(&L.d_blur, (size_t)ex->slab_fstride * B));
  ORBX_CUDA(ex, cudaMemsetAsync(L.d_in, 0, (size_t)ex->in_fstride * B, L.stream));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.slots, (size_t)P.slots_per_frame * B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.cand, (size_t)P.slots_per_frame * B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.lab, (size_t)P.slots_per_frame * B * 2));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.cell_count, (size_t)P.cells_per_frame * B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.lvl_kp, (size_t)P.kps_per_frame * B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.dst, (size_t)P.kps_per_frame * B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.lvl_n, (size_t)P.nlevels * B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.lvl_c, (size_t)P.nlevels * B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.ws.lvl_st, (size_t)P.nlevels * B * 4));
#ifdef ORBX_QT_PROF
  ORBX_CUDA(ex, cudaMalloc(&L.ws.qt_prof, (size_t)P.nlevels * B * 16 * 8));
#endif
  return ORBX_OK;
}

int alloc_lane_out(orbx_extractor* ex, OrbxLane& L, int cap) {
  if (L.out_cap >= cap && L.d_kps) return ORBX_OK;
  orbx::alloc_generation()++;
  ORBX_CUDA(ex, cudaStreamSynchronize(L.stream));
  cudaFree(L.d_kps);
  cudaFree(L.d_desc);
  cudaFree(L.d_n);
  cudaFree(L.d_mono);
  cudaFree(L.d_status);
  L.d_kps = nullptr;
  const size_t B = ex->max_batch;
  ORBX_CUDA(ex, cudaMalloc(&L.d_kps, (size_t)cap * B * sizeof(orbx_kp)));
  ORBX_CUDA(ex, cudaMalloc(&L.d_desc, (size_t)cap * B * ORBX_DESC_BYTES));
  ORBX_CUDA(ex, cudaMalloc(&L.d_n, B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.d_mono, B * 4));
  ORBX_CUDA(ex, cudaMalloc(&L.d_status, B * 4));
  L.out_cap = cap;
  return ORBX_OK;
}

// The whole extractor for `frames` frames whose level 0 is described by fs.{lvl0,pitch0,fstride0}, on lane ln.
int run_pipeline(orbx_extractor* ex, int ln, FrameSet fs, int frames, int lap0, int lap1, const OutSet& out,
                 cudaStream_t st, int f0 = 0) {
  const Plan& P = ex->plan;
  OrbxLane& L = ex->lane[ln];
  int rc = alloc_lane(ex, L);
  if (rc) return rc;
  fs.pyr = L.d_pyr;
  fs.blur = L.d_blur;
  fs.slab_fstride = ex->slab_fstride;
  bool prof = ex->profile;
  if (prof) {  // take 2 * kStages events from the pool: (start, end) of every stage
    while (ex->prof_events.size() < ex->prof_used + 2 * kStages) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { prof = false; break; }
      ex->prof_events.push_back(e);
    }
  }
  auto begin = [&](int stage, cudaStream_t s) {
    if (prof) cudaEventRecord(ex->prof_events[ex->prof_used + 2 * stage], s);
  };
  auto end = [&](int stage, cudaStream_t s) {
    if (prof) cudaEventRecord(ex->prof_events[ex->prof_used + 2 * stage + 1], s);
  };
  // One stream, stages back to back. (Forking the blur onto a second stream so that it overlaps FAST + quadtree was
  // measured SLOWER on B200: 7.01 vs 6.78 ms per 512-frame step — the blur's CTAs crowd out the latency-bound
  // quadtree warps and slow FAST, and nothing is gained because every stage already fills the chip.)
  static const bool skip_kernels = getenv("ORBX_DEBUG_SKIP_KERNELS") != nullptr;  // timing experiment only
  if (skip_kernels) {
    L.last_fs = fs;
    L.last_frames = frames;
    L.last_f0 = f0;
    L.last_call = ex->call_id;
    ex->last_lane = ln;
    return ORBX_OK;
  }
  static const bool blur_late = getenv("ORBX_BLUR_LATE") != nullptr;  // A/B experiment: blur after the quadtree
  begin(0, st);
  launch_pyramid(P, fs, ex->d_tab, frames, st);
  end(0, st);
  if (!blur_late) {
    begin(3, st);
    launch_blur(P, fs, frames, st);
    end(3, st);
  }
  begin(1, st);
  launch_fast(P, fs, L.ws, ex->ini_th, ex->min_th, frames, st);
  end(1, st);
  begin(2, st);
  launch_quadtree(P, L.ws, lap0, lap1, frames, st);
  end(2, st);
  if (blur_late) {
    begin(3, st);
    launch_blur(P, fs, frames, st);
    end(3, st);
  }
  begin(4, st);
  launch_describe(P, fs, L.ws, out, ex->d_pattern, frames, st);
  end(4, st);
  ORBX_CUDA(ex, cudaGetLastError());
  if (prof) ex->prof_used += 2 * kStages;
  L.last_fs = fs;
  L.last_frames = frames;
  L.last_f0 = f0;
  L.last_call = ex->call_id;
  ex->last_lane = ln;
  return ORBX_OK;
}

}  // namespace

namespace orbx {

int api_fail(orbx_extractor* ex, int code, const std::string& msg) {
  if (ex) ex->err = msg;
  else g_create_error = msg;
  return code;
}

void api_begin_call(orbx_extractor* ex) { ex->call_id++; }

int api_find_frame(const orbx_extractor* ex, int frame, int* lane, int* local) {
  if (!ex || !ex->planned || frame < 0) return ORBX_E_ARG;
  for (int ln = 0; ln < kLanes; ln++) {
    const OrbxLane& L = ex->lane[ln];
    if (L.last_call == ex->call_id && L.last_frames > 0 && frame >= L.last_f0 && frame < L.last_f0 + L.last_frames) {
      *lane = ln;
      *local = frame - L.last_f0;
      return ORBX_OK;
    }
  }
  return ORBX_E_ARG;
}

int api_ensure_plan(orbx_extractor* ex, int w, int h) {
  if (ex->planned && ex->plan.w == w && ex->plan.h == h) return ORBX_OK;
  ORBX_CUDA(ex, cudaSetDevice(ex->device));
  if (ex->planned) {
    int rc = sync_all_lanes(ex);
    if (rc) return rc;
    free_plan_buffers(ex);
  }
  Plan P;
  const int rc = make_plan(w, h, ex->nfeatures, ex->scale_factor, ex->nlevels, &P);
  if (rc == -2) return api_fail(ex, ORBX_E_SIZE, "image too small for the pyramid depth (a level has no 35-px cell)");
  if (rc == -3) return api_fail(ex, ORBX_E_SIZE, "image larger than 4096 px");
  if (rc != 0) return api_fail(ex, ORBX_E_ARG, "bad extractor parameters");
  for (int l = 0; l < P.nlevels; l++)
    if ((int64_t)P.lv[l].nCols * P.lv[l].nRows * P.lv[l].slot_cap >= (1 << 20))
      return api_fail(ex, ORBX_E_SIZE, "level too large: more than 2^20 candidate slots");
  ex->plan = P;
  ex->slab_fstride = (P.pyr_bytes_per_frame + 255) / 256 * 256;
  ex->in_pitch = round_up(w, 64);
  ex->in_fstride = (int64_t)ex->in_pitch * h;
  // resize tables
  std::vector<ResizeTab> tab(std::max(P.tab_entries, 1));
  {
    std::vector<int16_t> ofs, c0, c1;
    for (int l = 1; l < P.nlevels; l++) {
      const LevelPlan &D = P.lv[l], &S = P.lv[l - 1];
      for (int axis = 0; axis < 2; axis++) {
        const int ds = axis ? D.h : D.w, ss = axis ? S.h : S.w;
        ofs.resize(ds);
        c0.resize(ds);
        c1.resize(ds);
        axis_table(ss, ds, axis == 0, ofs.data(), c0.data(), c1.data());
        ResizeTab* t = tab.data() + (axis ? D.ytab_off : D.xtab_off);
        for (int d = 0; d < ds; d++) t[d] = ResizeTab{ofs[d], c0[d], c1[d], 0};
      }
    }
  }
  ORBX_CUDA(ex, cudaMalloc(&ex->d_tab, tab.size() * sizeof(ResizeTab)));
  ORBX_CUDA(ex, cudaMemcpy(ex->d_tab, tab.data(), tab.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice));
  ex->planned = true;
  return ORBX_OK;
}

int api_ensure_out(orbx_extractor* ex, int cap) {
  for (OrbxLane& L : ex->lane) {
    int rc = alloc_lane_out(ex, L, cap);
    if (rc) return rc;
  }
  return ORBX_OK;
}

thread_local cudaEvent_t g_trace_after_h2d = nullptr;

int api_upload_and_run(orbx_extractor* ex, int ln, const uint8_t* src, int nb, int width, int height, int stride,
                       int64_t frame_stride, int lap0, int lap1, cudaStream_t st, int f0, cudaStream_t copy_stream) {
  OrbxLane& L = ex->lane[ln];
  int rc = alloc_lane(ex, L);
  if (rc) return rc;
  const cudaStream_t cs = copy_stream ? copy_stream : st;
  FrameSet fs{};
  fs.lvl0 = L.d_in;
  if (frame_stride == (int64_t)stride * height && stride <= ex->in_pitch) {
    // densely packed frames: ONE contiguous copy, level 0 keeps the caller's pitch (a 2-D copy of 752-byte rows runs at
    // a fraction of the PCIe rate)
    // ORBX_DEBUG_SKIP_H2D: timing experiment only (what the pipelined calls cost without the upload); results are then
    // those of whatever the lane held before
    static const bool skip_h2d = getenv("ORBX_DEBUG_SKIP_H2D") != nullptr;
    if (!skip_h2d)
      ORBX_CUDA(ex, cudaMemcpyAsync(L.d_in, src, (size_t)frame_stride * nb, cudaMemcpyHostToDevice, cs));
    if (g_trace_after_h2d) cudaEventRecord(g_trace_after_h2d, cs);  // ORBX_TRACE=2 timeline of the pipelined calls
    fs.pitch0 = stride;
    fs.fstride0 = frame_stride;
  } else {
    for (int f = 0; f < nb; f++)
      ORBX_CUDA(ex, cudaMemcpy2DAsync(L.d_in + f * ex->in_fstride, ex->in_pitch, src + f * frame_stride, stride,
                                      width, height, cudaMemcpyHostToDevice, cs));
    fs.pitch0 = ex->in_pitch;
    fs.fstride0 = ex->in_fstride;
  }
  if (copy_stream) {
    ORBX_CUDA(ex, cudaEventRecord(L.up, copy_stream));
    ORBX_CUDA(ex, cudaStreamWaitEvent(st, L.up, 0));
  }
  OutSet out{L.d_kps, L.d_desc, L.d_n, L.d_mono, L.d_status, L.out_cap};
  return run_pipeline(ex, ln, fs, nb, lap0, lap1, out, st, f0);
}

int api_download(orbx_extractor* ex, int ln, int nb, orbx_kp* kps, uint8_t* desc, int cap, cudaStream_t st) {
  OrbxLane& L = ex->lane[ln];
  const int B = ex->max_batch;
  ORBX_CUDA(ex, cudaMemcpyAsync(L.h_small, L.d_n, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(ex, cudaMemcpyAsync(L.h_small + B, L.d_mono, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(ex, cudaMemcpyAsync(L.h_small + 2 * B, L.d_status, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
  // rows: device arrays are [nb][out_cap], the caller's are [..][cap]
  if (cap == L.out_cap) {
    ORBX_CUDA(ex, cudaMemcpyAsync(kps, L.d_kps, (size_t)nb * cap * sizeof(orbx_kp), cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(ex, cudaMemcpyAsync(desc, L.d_desc, (size_t)nb * cap * ORBX_DESC_BYTES, cudaMemcpyDeviceToHost, st));
  } else {
    const int rows = std::min(cap, L.out_cap);
    ORBX_CUDA(ex, cudaMemcpy2DAsync(kps, (size_t)cap * sizeof(orbx_kp), L.d_kps, (size_t)L.out_cap * sizeof(orbx_kp),
                                    (size_t)rows * sizeof(orbx_kp), nb, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(ex, cudaMemcpy2DAsync(desc, (size_t)cap * ORBX_DESC_BYTES, L.d_desc,
                                    (size_t)L.out_cap * ORBX_DESC_BYTES, (size_t)rows * ORBX_DESC_BYTES, nb,
                                    cudaMemcpyDeviceToHost, st));
  }
  return ORBX_OK;
}

}  // namespace orbx

extern "C" {

int orbx_extractor_create(orbx_extractor** out, int device, int nfeatures, float scale_factor, int nlevels,
                          int ini_th_fast, int min_th_fast, int max_batch) {
  if (!out) return api_fail(nullptr, ORBX_E_ARG, "out == NULL");
  *out = nullptr;
  if (nfeatures < 1 || nlevels < 1 || nlevels > kMaxLevels || !(scale_factor > 1.0f) || max_batch < 1 ||
      ini_th_fast < 0 || ini_th_fast > 255 || min_th_fast < 0 || min_th_fast > 255)
    return api_fail(nullptr, ORBX_E_ARG, "bad extractor parameters");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return api_fail(nullptr, ORBX_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return api_fail(nullptr, ORBX_E_ARG, "bad device ordinal");
  orbx_extractor* ex = new orbx_extractor;
  ex->device = device;
  ex->nfeatures = nfeatures;
  ex->scale_factor = scale_factor;
  ex->nlevels = nlevels;
  ex->ini_th = ini_th_fast;
  ex->min_th = min_th_fast;
  ex->max_batch = max_batch;
  auto bail = [&](const char* what, cudaError_t ce) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
    orbx_extractor_destroy(ex);
    return ORBX_E_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  for (OrbxLane& L : ex->lane) {
    if ((e = cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking)) != cudaSuccess)
      return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming)) != cudaSuccess)
      return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&L.up, cudaEventDisableTiming)) != cudaSuccess)
      return bail("cudaEventCreate", e);
    if ((e = cudaHostAlloc(&L.h_small, (size_t)3 * max_batch * 4, cudaHostAllocDefault)) != cudaSuccess)
      return bail("cudaHostAlloc", e);
  }
  if ((e = cudaMalloc(&ex->d_pattern, sizeof(kPatternHost))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemcpy(ex->d_pattern, kPatternHost, sizeof(kPatternHost), cudaMemcpyHostToDevice)) != cudaSuccess)
    return bail("cudaMemcpy", e);
  *out = ex;
  return ORBX_OK;
}

void orbx_extractor_destroy(orbx_extractor* ex) {
  if (!ex) return;
  cudaSetDevice(ex->device);
  for (OrbxLane& L : ex->lane)
    if (L.stream) cudaStreamSynchronize(L.stream);
  free_plan_buffers(ex);
  free_out_buffers(ex);
  cudaFree(ex->d_pattern);
  for (auto& ev : ex->prof_events) cudaEventDestroy(ev);
  for (OrbxLane& L : ex->lane) {
    if (L.h_small) cudaFreeHost(L.h_small);
    if (L.done) cudaEventDestroy(L.done);
    if (L.up) cudaEventDestroy(L.up);
    if (L.stream) cudaStreamDestroy(L.stream);
  }
  delete ex;
}

const char* orbx_last_error(const orbx_extractor* ex) { return ex ? ex->err.c_str() : g_create_error.c_str(); }

int orbx_extractor_levels(const orbx_extractor* ex) { return ex ? ex->nlevels : ORBX_E_ARG; }

int orbx_extractor_tables(const orbx_extractor* ex, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int32_t* features_per_level) {
  if (!ex) return ORBX_E_ARG;
  // the tables do not depend on the image size (and exist for level counts no image of <= 4096 px could carry)
  float sf[kMaxLevels], s2[kMaxLevels];
  int quota[kMaxLevels];
  make_tables(ex->nfeatures, ex->scale_factor, ex->nlevels, sf, s2, quota);
  for (int l = 0; l < ex->nlevels; l++) {
    if (scale) scale[l] = sf[l];
    if (inv_scale) inv_scale[l] = 1.0f / sf[l];
    if (sigma2) sigma2[l] = s2[l];
    if (inv_sigma2) inv_sigma2[l] = 1.0f / s2[l];
    if (features_per_level) features_per_level[l] = quota[l];
  }
  return ORBX_OK;
}

int orbx_extractor_capacity(const orbx_extractor* ex) {
  if (!ex) return ORBX_E_ARG;
  // every level may overshoot its quota by 3 (or return the 4*nIni nodes of the first split when the quota is tiny)
  return ex->nfeatures + 16 * ex->nlevels;
}

int orbx_extract_batch_device(orbx_extractor* ex, int n_frames, const uint8_t* d_images, int width, int height,
                              int stride, int64_t frame_stride, int lap0, int lap1, orbx_kp* d_kps, uint8_t* d_desc,
                              int cap, int32_t* d_n, int32_t* d_mono_index, int32_t* d_status, void* cuda_stream) {
  if (!ex) return ORBX_E_ARG;
  if (!d_images || width <= 0 || height <= 0 || n_frames <= 0) return api_fail(ex, ORBX_E_EMPTY, "empty image");
  if (n_frames > ex->max_batch) return api_fail(ex, ORBX_E_ARG, "n_frames > max_batch");
  if (stride < width || !d_kps || !d_desc || !d_n || !d_mono_index || !d_status || cap < 1)
    return api_fail(ex, ORBX_E_ARG, "bad argument");
  ORBX_CUDA(ex, cudaSetDevice(ex->device));
  int rc = api_ensure_plan(ex, width, height);
  if (rc) return rc;
  api_begin_call(ex);
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ex->lane[0].stream;
  FrameSet fs{};
  fs.lvl0 = d_images;
  fs.pitch0 = stride;
  fs.fstride0 = frame_stride;
  OutSet out{d_kps, d_desc, d_n, d_mono_index, d_status, cap};
  return run_pipeline(ex, 0, fs, n_frames, lap0, lap1, out, st);
}

int orbx_extract_batch(orbx_extractor* ex, int n_frames, const uint8_t* images, int width, int height, int stride,
                       int64_t frame_stride, int lap0, int lap1, orbx_kp* kps, uint8_t* desc, int cap,
                       int32_t* n_out, int32_t* mono_index) {
  if (!ex) return ORBX_E_ARG;
  if (!images || width <= 0 || height <= 0 || n_frames <= 0) return api_fail(ex, ORBX_E_EMPTY, "empty image");
  if (stride < width || !kps || !desc || !n_out || cap < 1) return api_fail(ex, ORBX_E_ARG, "bad argument");
  ORBX_CUDA(ex, cudaSetDevice(ex->device));
  int rc = api_ensure_plan(ex, width, height);
  if (rc) return rc;
  rc = api_ensure_out(ex, cap);
  if (rc) return rc;
  api_begin_call(ex);
  const int B = ex->max_batch;
  int first_err = ORBX_OK;
  int pending_f0[kLanes], pending_nb[kLanes];
  for (int i = 0; i < kLanes; i++) pending_nb[i] = 0;
  // collect the small per-frame results of the group a lane finished
  auto retire = [&](int ln) -> int {
    if (pending_nb[ln] == 0) return ORBX_OK;
    OrbxLane& L = ex->lane[ln];
    ORBX_CUDA(ex, cudaEventSynchronize(L.done));
    for (int f = 0; f < pending_nb[ln]; f++) {
      const int g = pending_f0[ln] + f;
      n_out[g] = L.h_small[f];
      if (mono_index) mono_index[g] = L.h_small[B + f];
      if ((L.h_small[2 * B + f] != 0 || n_out[g] > cap) && first_err == ORBX_OK) first_err = ORBX_E_CAPACITY;
    }
    pending_nb[ln] = 0;
    return ORBX_OK;
  };
  int group = 0;
  for (int f0 = 0; f0 < n_frames; f0 += B, group++) {
    const int nb = std::min(B, n_frames - f0);
    const int ln = group % kLanes;
    OrbxLane& L = ex->lane[ln];
    if ((rc = retire(ln)) != 0) return rc;
    rc = api_upload_and_run(ex, ln, images + (int64_t)f0 * frame_stride, nb, width, height, stride, frame_stride,
                            lap0, lap1, L.stream, f0);
    if (rc) return rc;
    rc = api_download(ex, ln, nb, kps + (int64_t)f0 * cap, desc + (int64_t)f0 * cap * ORBX_DESC_BYTES, cap, L.stream);
    if (rc) return rc;
    ORBX_CUDA(ex, cudaEventRecord(L.done, L.stream));
    pending_f0[ln] = f0;
    pending_nb[ln] = nb;
  }
  for (int ln = 0; ln < kLanes; ln++)
    if ((rc = retire(ln)) != 0) return rc;
  if (first_err) return api_fail(ex, first_err, "output capacity too small for at least one frame");
  return ORBX_OK;
}

int orbx_extract_batch_multi(int n_devices, orbx_extractor* const* ex, int n_frames, const uint8_t* images, int width,
                             int height, int stride, int64_t frame_stride, int lap0, int lap1, orbx_kp* kps,
                             uint8_t* desc, int cap, int32_t* n_out, int32_t* mono_index) {
  if (n_devices < 1 || !ex) return ORBX_E_ARG;
  for (int d = 0; d < n_devices; d++)
    if (!ex[d]) return ORBX_E_ARG;
  if (n_frames <= 0 || !images) return api_fail(ex[0], ORBX_E_EMPTY, "empty image");
  std::vector<int> rc(n_devices, ORBX_OK);
  std::vector<std::thread> th;
  const int base = n_frames / n_devices, extra = n_frames % n_devices;
  int f0 = 0;
  for (int d = 0; d < n_devices; d++) {
    const int nb = base + (d < extra ? 1 : 0);
    if (nb > 0)
      th.emplace_back([=, &rc] {
        rc[d] = orbx_extract_batch(ex[d], nb, images + (int64_t)f0 * frame_stride, width, height, stride, frame_stride, lap0,
                                   lap1, kps + (int64_t)f0 * cap, desc + (int64_t)f0 * cap * ORBX_DESC_BYTES, cap,
                                   n_out + f0, mono_index ? mono_index + f0 : nullptr);
      });
    f0 += nb;
  }
  for (auto& t : th) t.join();
  for (int d = 0; d < n_devices; d++)
    if (rc[d] != ORBX_OK) return rc[d];
  return ORBX_OK;
}

int orbx_extract(orbx_extractor* ex, const uint8_t* image, int width, int height, int stride, int lap0, int lap1,
                 orbx_kp* kps, uint8_t* desc, int cap, int32_t* n_out, int32_t* mono_index) {
  int32_t n = 0, mono = 0;
  const int rc = orbx_extract_batch(ex, 1, image, width, height, stride, (int64_t)stride * height, lap0, lap1, kps,
                                    desc, cap, &n, &mono);
  if (n_out) *n_out = n;
  if (mono_index) *mono_index = mono;
  return rc;
}

int orbx_level_size(const orbx_extractor* ex, int level, int* width, int* height) {
  if (!ex || !ex->planned || level < 0 || level >= ex->nlevels) return ORBX_E_ARG;
  if (width) *width = ex->plan.lv[level].w;
  if (height) *height = ex->plan.lv[level].h;
  return ORBX_OK;
}

int orbx_debug_level(orbx_extractor* ex, int frame, int level, int which, uint8_t* dst, int dst_stride) {
  int fl = 0, lf = 0;
  if (!ex || !ex->planned || level < 0 || level >= ex->nlevels || !dst || api_find_frame(ex, frame, &fl, &lf) != ORBX_OK)
    return ex ? api_fail(ex, ORBX_E_ARG, "bad argument, or the frame is no longer resident (only the last kLanes groups of a call are)") : ORBX_E_ARG;
  frame = lf;
  ORBX_CUDA(ex, cudaSetDevice(ex->device));
  const LevelPlan& L = ex->plan.lv[level];
  const FrameSet& fs = ex->lane[fl].last_fs;
  const uint8_t* src;
  int pitch;
  if (which == 1) {
    src = fs.blur + (int64_t)frame * fs.slab_fstride + L.img_off;
    pitch = L.pitch;
  } else if (level == 0) {
    src = fs.lvl0 + (int64_t)frame * fs.fstride0;
    pitch = fs.pitch0;
  } else {
    src = fs.pyr + (int64_t)frame * fs.slab_fstride + L.img_off;
    pitch = L.pitch;
  }
  ORBX_CUDA(ex, cudaDeviceSynchronize());
  ORBX_CUDA(ex, cudaMemcpy2D(dst, dst_stride, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost));
  return ORBX_OK;
}

int orbx_download_pyramid(orbx_extractor* ex, int frame, int level, uint8_t* dst, int dst_stride) {
  if (!ex || !ex->planned || level < 0 || level >= ex->nlevels || !dst) return ORBX_E_ARG;
  const LevelPlan& L = ex->plan.lv[level];
  if (dst_stride < L.w + 2 * kEdge) return api_fail(ex, ORBX_E_ARG, "dst_stride < w + 38");
  // interior first, then the reflect-101 frame is a pure copy of interior pixels (cv::copyMakeBorder, :1129-1143)
  uint8_t* roi = dst + (size_t)kEdge * dst_stride + kEdge;
  int rc = orbx_debug_level(ex, frame, level, 0, roi, dst_stride);
  if (rc) return rc;
  for (int y = 0; y < L.h; y++) {
    uint8_t* row = roi + (size_t)y * dst_stride;
    for (int x = 1; x <= kEdge; x++) {
      row[-x] = row[reflect101(-x, L.w)];
      row[L.w - 1 + x] = row[reflect101(L.w - 1 + x, L.w)];
    }
  }
  for (int y = 1; y <= kEdge; y++) {
    memcpy(dst + (size_t)(kEdge - y) * dst_stride, dst + (size_t)(kEdge + reflect101(-y, L.h)) * dst_stride,
           L.w + 2 * kEdge);
    memcpy(dst + (size_t)(kEdge + L.h - 1 + y) * dst_stride,
           dst + (size_t)(kEdge + reflect101(L.h - 1 + y, L.h)) * dst_stride, L.w + 2 * kEdge);
  }
  return ORBX_OK;
}

static void unpack_kp(uint32_t cw, int add, int level, float size, orbx_kp* k) {
  k->x = (float)(cand_x(cw) + add);
  k->y = (float)(cand_y(cw) + add);
  k->size = size;
  k->angle = -1.f;
  k->response = (float)cand_s(cw);
  k->octave = level;
  k->class_id = -1;
}

int orbx_debug_candidates(orbx_extractor* ex, int frame, int level, orbx_kp* out, int cap) {
  int fl = 0, lf = 0;
  if (!ex || !ex->planned || level < 0 || level >= ex->nlevels || api_find_frame(ex, frame, &fl, &lf) != ORBX_OK)
    return ORBX_E_ARG;
  frame = lf;
  ORBX_CUDA(ex, cudaSetDevice(ex->device));
  ORBX_CUDA(ex, cudaDeviceSynchronize());
  // the candidates of a level in the order the reference appends them (:905-958) = the per-cell slots written by
  // k_fast, cells in row-major order (pure data movement: the quadtree kernel builds the same array in shared memory)
  const Plan& P = ex->plan;
  const LevelPlan& L = P.lv[level];
  const OrbxLane& ln = ex->lane[fl];
  const int ncell = L.nCols * L.nRows;
  std::vector<int32_t> cnt(ncell);
  std::vector<uint32_t> slots((size_t)ncell * L.slot_cap);
  ORBX_CUDA(ex, cudaMemcpy(cnt.data(), ln.ws.cell_count + (int64_t)frame * P.cells_per_frame + L.cell_base,
                           (size_t)ncell * 4, cudaMemcpyDeviceToHost));
  ORBX_CUDA(ex, cudaMemcpy(slots.data(), ln.ws.slots + (int64_t)frame * P.slots_per_frame + L.slot_base,
                           slots.size() * 4, cudaMemcpyDeviceToHost));
  int C = 0;
  for (int c = 0; c < ncell; c++)
    for (int k = 0; k < cnt[c]; k++, C++)
      if (out && C < cap) unpack_kp(slots[(size_t)c * L.slot_cap + k], 0, 0, 7.f, out + C);
  return C;
}

int orbx_debug_level_keypoints(orbx_extractor* ex, int frame, int level, orbx_kp* out, int cap) {
  int fl = 0, lf = 0;
  if (!ex || !ex->planned || level < 0 || level >= ex->nlevels || api_find_frame(ex, frame, &fl, &lf) != ORBX_OK)
    return ORBX_E_ARG;
  frame = lf;
  ORBX_CUDA(ex, cudaSetDevice(ex->device));
  ORBX_CUDA(ex, cudaDeviceSynchronize());
  const Plan& P = ex->plan;
  int32_t C = 0;
  ORBX_CUDA(ex, cudaMemcpy(&C, ex->lane[fl].ws.lvl_n + frame * P.nlevels + level, 4, cudaMemcpyDeviceToHost));
  const int n = std::min(C, cap);
  std::vector<uint32_t> buf(std::max(n, 1));
  ORBX_CUDA(ex, cudaMemcpy(buf.data(), ex->lane[fl].ws.lvl_kp + (int64_t)frame * P.kps_per_frame + P.lv[level].kp_base,
                           (size_t)n * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n && out; i++) unpack_kp(buf[i], kMinBorder, level, (float)P.lv[level].patch, out + i);
  return C;
}

// Cycle accounting of k_quadtree for (frame, level) of the last call: out[16] (see orbx_quadtree.h). Only in builds
// with -DORBX_QT_PROF; otherwise ORBX_E_ARG. Development aid, not part of the public header.
int orbx_debug_qt_profile(orbx_extractor* ex, int frame, int level, long long* out) {
  if (!ex || !ex->planned || !out) return ORBX_E_ARG;
  int fl = 0, lf = 0;
  if (api_find_frame(ex, frame, &fl, &lf) != ORBX_OK) return ORBX_E_ARG;
  frame = lf;
  const OrbxLane& ln = ex->lane[fl];
  if (!ln.ws.qt_prof || level < 0 || level >= ex->nlevels) return ORBX_E_ARG;
  ORBX_CUDA(ex, cudaSetDevice(ex->device));
  ORBX_CUDA(ex, cudaDeviceSynchronize());
  ORBX_CUDA(ex, cudaMemcpy(out, ln.ws.qt_prof + ((int64_t)frame * ex->nlevels + level) * 16, 16 * 8,
                           cudaMemcpyDeviceToHost));
  return ORBX_OK;
}

int orbx_profile_enable(orbx_extractor* ex, int on) {
  if (!ex) return ORBX_E_ARG;
  ex->profile = on != 0;
  return ORBX_OK;
}

int orbx_profile_read(orbx_extractor* ex, float* ms, int32_t* launches, int reset) {
  if (!ex) return ORBX_E_ARG;
  ORBX_CUDA(ex, cudaSetDevice(ex->device));
  for (size_t r = 0; r + 2 * kStages <= ex->prof_used; r += 2 * kStages) {
    for (int s = 0; s < kStages; s++) {
      float t = 0;
      ORBX_CUDA(ex, cudaEventSynchronize(ex->prof_events[r + 2 * s + 1]));
      ORBX_CUDA(ex, cudaEventElapsedTime(&t, ex->prof_events[r + 2 * s], ex->prof_events[r + 2 * s + 1]));
      ex->prof_ms[s] += t;
      ex->prof_launches[s] += 1;
    }
  }
  ex->prof_used = 0;
  for (int s = 0; s < kStages; s++) {
    if (ms) ms[s] = ex->prof_ms[s];
    if (launches) launches[s] = ex->prof_launches[s];
    if (reset) {
      ex->prof_ms[s] = 0;
      ex->prof_launches[s] = 0;
    }
  }
  return ORBX_OK;
}

int orbx_remap_linear_device(int device, int n_frames, const uint8_t* d_src, int src_width, int src_height,
                             int src_stride, int64_t src_frame_stride, const float* d_mapx, const float* d_mapy,
                             int dst_width, int dst_height, uint8_t* d_dst, int dst_stride, int64_t dst_frame_stride,
                             void* cuda_stream) {
  if (!d_src || !d_dst || !d_mapx || !d_mapy || src_width <= 0 || src_height <= 0 || dst_width <= 0 ||
      dst_height <= 0 || n_frames <= 0 || src_stride < src_width || dst_stride < dst_width)
    return ORBX_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return ORBX_E_CUDA;
  orbx::launch_remap_linear(d_src, src_width, src_height, src_stride, src_frame_stride, d_mapx, d_mapy, dst_width,
                            dst_height, d_dst, dst_stride, dst_frame_stride, n_frames, (cudaStream_t)cuda_stream);
  return cudaGetLastError() == cudaSuccess ? ORBX_OK : ORBX_E_CUDA;
}

int orbx_remap_linear(int device, const uint8_t* src, int src_width, int src_height, int src_stride, const float* mapx,
                      const float* mapy, int dst_width, int dst_height, uint8_t* dst, int dst_stride) {
  if (!src || !dst || !mapx || !mapy || src_width <= 0 || src_height <= 0 || dst_width <= 0 || dst_height <= 0 ||
      src_stride < src_width || dst_stride < dst_width)
    return ORBX_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return ORBX_E_CUDA;
  uint8_t *d_src = nullptr, *d_dst = nullptr;
  float *d_mx = nullptr, *d_my = nullptr;
  const size_t sb = (size_t)src_stride * src_height, db = (size_t)dst_stride * dst_height,
               mb = (size_t)dst_width * dst_height * sizeof(float);
  int rc = ORBX_E_CUDA;
  if (cudaMalloc(&d_src, sb) == cudaSuccess && cudaMalloc(&d_dst, db) == cudaSuccess &&
      cudaMalloc(&d_mx, mb) == cudaSuccess && cudaMalloc(&d_my, mb) == cudaSuccess &&
      cudaMemcpy(d_src, src, sb, cudaMemcpyHostToDevice) == cudaSuccess &&
      cudaMemcpy(d_mx, mapx, mb, cudaMemcpyHostToDevice) == cudaSuccess &&
      cudaMemcpy(d_my, mapy, mb, cudaMemcpyHostToDevice) == cudaSuccess) {
    rc = orbx_remap_linear_device(device, 1, d_src, src_width, src_height, src_stride, 0, d_mx, d_my, dst_width,
                                  dst_height, d_dst, dst_stride, 0, nullptr);
    if (rc == ORBX_OK &&
        cudaMemcpy2D(dst, dst_stride, d_dst, dst_stride, dst_width, dst_height, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = ORBX_E_CUDA;
  }
  cudaFree(d_src);
  cudaFree(d_dst);
  cudaFree(d_mx);
  cudaFree(d_my);
  return rc;
}

int orbx_cvt_gray_device(int device, int n_frames, const uint8_t* d_src, int width, int height, int src_stride,
                         int64_t src_frame_stride, int channels, int rgb, uint8_t* d_dst, int dst_stride,
                         int64_t dst_frame_stride, void* cuda_stream) {
  if (!d_src || !d_dst || width <= 0 || height <= 0 || n_frames <= 0 || (channels != 3 && channels != 4) ||
      src_stride < width * channels || dst_stride < width)
    return ORBX_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return ORBX_E_CUDA;
  if (orbx::launch_cvt_gray(d_src, width, height, src_stride, src_frame_stride, channels, rgb, d_dst, dst_stride,
                            dst_frame_stride, n_frames, (cudaStream_t)cuda_stream) != 0)
    return ORBX_E_ARG;
  return cudaGetLastError() == cudaSuccess ? ORBX_OK : ORBX_E_CUDA;
}

int orbx_cvt_gray(int device, const uint8_t* src, int width, int height, int src_stride, int channels, int rgb,
                  uint8_t* dst, int dst_stride) {
  if (!src || !dst || width <= 0 || height <= 0 || (channels != 3 && channels != 4) ||
      src_stride < width * channels || dst_stride < width)
    return ORBX_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return ORBX_E_CUDA;
  uint8_t *d_src = nullptr, *d_dst = nullptr;
  const size_t sb = (size_t)src_stride * height, db = (size_t)dst_stride * height;
  int rc = ORBX_E_CUDA;
  if (cudaMalloc(&d_src, sb) == cudaSuccess && cudaMalloc(&d_dst, db) == cudaSuccess &&
      cudaMemcpy(d_src, src, sb, cudaMemcpyHostToDevice) == cudaSuccess &&
      cudaMemset(d_dst, 0, db) == cudaSuccess) {
    rc = orbx_cvt_gray_device(device, 1, d_src, width, height, src_stride, 0, channels, rgb, d_dst, dst_stride, 0, nullptr);
    if (rc == ORBX_OK && cudaMemcpy2D(dst, dst_stride, d_dst, dst_stride, width, height, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = ORBX_E_CUDA;
  }
  cudaFree(d_src);
  cudaFree(d_dst);
  return rc;
}

int orbx_kernel_launches(const orbx_extractor* ex) {
  if (!ex || !ex->planned) return 0;
  return (ex->plan.nlevels - 1) + orbx::fast_launch_count(ex->plan) + 3;  // resize per level, FAST, quadtree, blur, describe
}

void* orbx_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void orbx_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
