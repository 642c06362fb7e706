// orbm_handle.h — the matcher context behind include/orbm.h, shared by the translation units of the matcher ABI
// (orbm_api.cu, orbm_track_api.cu).
#ifndef ORBM_HANDLE_H_
#define ORBM_HANDLE_H_

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/orbm.h"
#include "orbx_handle.h"
#include "orbx_match.cuh"

namespace orbm_detail {

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    orbx::alloc_generation()++;  // recorded graphs that point into this buffer must not be replayed
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = std::max(bytes, (size_t)4096) * 5 / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) orbx::alloc_generation()++;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};
enum { kBufs = 40, kTrackBufs = 12 };

// A small stereo call (one pipeline group) issues ~60 kernels and ~20 copies / event operations whose host-side cost
// (~0.8 ms) exceeds the GPU's (~0.5 ms for one 640x480 pair). When the SAME call comes back — same buffers, same sizes,
// same parameters: the loop of an online front-end — its work is recorded once as a CUDA graph and replayed.
struct StereoGraphKey {  // compared with memcmp: always memset before filling
  const void* ptr[28];
  long long iv[12];
  float fv[12];
  unsigned long long gen;  // orbx::alloc_generation() the call ran under
};
struct StereoGraphSlot {
  StereoGraphKey key;
  cudaGraphExec_t exec = nullptr;
  bool used = false;
  unsigned long long stamp = 0;
  orbx::FrameSet fs_l{}, fs_r{};  // host-side bookkeeping of the recorded run (OrbxLane::last_fs)
};
enum { kStereoGraphSlots = 4 };
}  // namespace orbm_detail
using orbm_detail::DevBuf;
using orbm_detail::kBufs;
using orbm_detail::kTrackBufs;

struct orbm_matcher {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  DevBuf buf[kBufs];
  int next_buf = 0;
  // per-lane device outputs of orbm_stereo_frames_batch: u_right, depth, sad [B][cap], n_matched [B] (+ pinned copy)
  DevBuf lane_buf[kLanes][4];
  int32_t* lane_h_nm[kLanes] = {};
  int lane_h_cap[kLanes] = {};
  // small host arrays of one call are packed into one pinned block and cross PCIe in ONE copy (a frame / keyframe
  // view is 8-10 arrays: 20 separate pageable copies cost more than the kernels of a guided search)
  uint8_t* h_stage = nullptr;
  uint8_t* d_stage = nullptr;
  size_t stage_used = 0, stage_flushed = 0;
  // the vocabulary tree of orbm_set_vocabulary (device resident across calls)
  DevBuf voc_buf[5];
  orbx::DevVocabulary voc{};
  // +-1 byte expansions of the query / train descriptors for the tensor-core knn2 (256 B per row)
  DevBuf tc_buf[2];
  // batched local-map tracking search (orbm_track_api.cu): scratch per pipeline lane (+1 for the device-resident call),
  // the device copy of the local maps of the host-facing call, pinned per-lane result words
  DevBuf track_buf[kLanes + 1][kTrackBufs];
  DevBuf track_map[8];
  cudaEvent_t track_map_ready = nullptr;
  // every H2D copy of a pipelined stereo call goes through this one stream: the copy engine then serves them in issue
  // order (with one stream per lane it served the channels round robin and the FIRST group's images arrived last, which
  // kept its lane busy and the copy engine idle for ~2.8 ms per call; ORBX_TRACE=2 timeline)
  cudaStream_t up_stream = nullptr;
  cudaEvent_t up_small[8] = {};  // per lane: the tracking stage's small inputs are on the device
  int32_t* lane_h_track[kLanes] = {};
  int lane_h_track_cap[kLanes] = {};
  // recorded small stereo calls (see StereoGraphKey)
  orbm_detail::StereoGraphSlot sgraph[orbm_detail::kStereoGraphSlots];
  unsigned long long sgraph_clock = 0;
  bool sgraph_off = false;          // set when a recording failed: the handle stays on the direct path
  bool sgraph_recording = false;    // a stream capture is open (stereo_frames_body)
  cudaEvent_t sgraph_fork = nullptr;
  unsigned long long sgraph_launches = 0;
};

namespace orbm_detail {

// ORBM_KNN2_TC=0 keeps every knn2 call on the POPC kernel (A/B runs, and the parity test of one path against the other)
inline bool knn2_tc_enabled() {
  const char* v = getenv("ORBM_KNN2_TC");
  return !(v && v[0] == '0');
}

std::string& create_error();  // thread-local text of the last failed orbm_create
inline int mfail(orbm_matcher* m, int code, const std::string& msg) {
  if (m) m->err = msg;
  else create_error() = msg;
  return code;
}

#define ORBM_CUDA(m, call)                                                                \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return mfail(m, ORBX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
  } while (0)

// One call = a sequence of scratch allocations in fixed order; buffers are reused across calls by position.
constexpr size_t kStageBytes = 8u << 20, kStageMaxItem = 512u << 10;

struct Arena {
  orbm_matcher* m;
  cudaError_t err = cudaSuccess;
  explicit Arena(orbm_matcher* mm) : m(mm) {
    m->next_buf = 0;
    m->stage_used = m->stage_flushed = 0;
    if (!m->h_stage) {
      if (cudaHostAlloc(reinterpret_cast<void**>(&m->h_stage), kStageBytes, cudaHostAllocDefault) != cudaSuccess ||
          cudaMalloc(reinterpret_cast<void**>(&m->d_stage), kStageBytes) != cudaSuccess) {
        if (m->h_stage) cudaFreeHost(m->h_stage);
        m->h_stage = nullptr;  // staging is an optimisation: fall back to one copy per array
        cudaGetLastError();
      }
    }
  }
  template <typename T>
  T* alloc(size_t count) {
    if (m->next_buf >= kBufs) {
      err = cudaErrorMemoryAllocation;
      return nullptr;
    }
    DevBuf& b = m->buf[m->next_buf++];
    cudaError_t e = b.reserve(std::max(count, (size_t)1) * sizeof(T));
    if (e != cudaSuccess) err = e;
    return reinterpret_cast<T*>(b.p);
  }
  template <typename T>
  T* upload(const T* host, size_t count) {
    const size_t bytes = count * sizeof(T);
    if (host && bytes && bytes <= kStageMaxItem && m->h_stage && m->stage_used + bytes <= kStageBytes) {
      const size_t off = m->stage_used;
      memcpy(m->h_stage + off, host, bytes);
      m->stage_used = (off + bytes + 255) & ~(size_t)255;
      return reinterpret_cast<T*>(m->d_stage + off);
    }
    T* d = alloc<T>(count);
    if (d && host && count) {
      cudaError_t e = cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, m->stream);
      if (e != cudaSuccess) err = e;
    }
    return d;
  }
  // Sends what upload() packed since the last call (one H2D copy on the matcher's stream) and reports the first error
  // of the arena. Every entry point calls it after its last upload and before its first kernel.
  cudaError_t sync_uploads() {
    if (m->stage_used > m->stage_flushed) {
      cudaError_t e = cudaMemcpyAsync(m->d_stage + m->stage_flushed, m->h_stage + m->stage_flushed,
                                      m->stage_used - m->stage_flushed, cudaMemcpyHostToDevice, m->stream);
      if (e != cudaSuccess && err == cudaSuccess) err = e;
      m->stage_flushed = m->stage_used;
    }
    return err;
  }
};

// orbm_track_api.cu: set-up shared by the device-resident tracking search and the host-facing pipelined call
int track_prepare(orbm_matcher* m, int slot, orbx::TrackArgs* A);
int track_fill_params(orbm_matcher* m, const orbx_extractor* ex, const orbx_local_map* maps, const orbx_track_params* prm,
                      int cap, orbx::TrackArgs* A);
void track_set_map(const orbx_local_map* device_map, orbx::TrackArgs* A);

}  // namespace orbm_detail
using namespace orbm_detail;

#endif
