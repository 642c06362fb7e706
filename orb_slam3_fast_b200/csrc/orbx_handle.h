// orbx_handle.h — the extractor handle, shared by the extractor ABI (orbx_api.cu) and the matcher ABI (orbm_api.cu:
// Frame::ComputeStereoMatches reads both extractors' pyramids, src/Frame.cc:927,1011,1029).
//
// A handle owns kLanes independent copies of the per-batch device state ("lanes"), each with its own stream. The
// host-facing batched calls cut the frames into groups of max_batch and rotate through the lanes, so the H2D copy of
// group i+1 and the D2H copy of group i-1 overlap the kernels of group i.
#ifndef ORBX_HANDLE_H_
#define ORBX_HANDLE_H_

#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "orbx_kernels.cuh"

// 8 lanes: the pipelined host-facing calls are bound by the H2D copies once the kernels are fast enough, and with 4
// lanes the copy engine idled ~1.8 ms per call while the host waited for the first group to retire (ORBX_TRACE=2 timeline)
enum { kStages = 5, kLanes = 8 };

struct OrbxLane {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  cudaEvent_t up = nullptr;  // recorded on the caller's upload stream after the lane's H2D copy (api_upload_and_run)
  // geometry-dependent device state
  uint8_t *d_in = nullptr, *d_pyr = nullptr, *d_blur = nullptr;
  orbx::WorkSet ws{};
  // device outputs of the host-facing calls, [max_batch][out_cap]
  int out_cap = 0;
  orbx_kp* d_kps = nullptr;
  uint8_t* d_desc = nullptr;
  int32_t *d_n = nullptr, *d_mono = nullptr, *d_status = nullptr;
  int32_t* h_small = nullptr;  // pinned [3][max_batch]: n, mono, status
  // what the lane holds since its last run (for downloads / stereo matching)
  orbx::FrameSet last_fs{};
  int last_frames = 0;
  int last_f0 = 0;             // global index (inside the call) of the lane's first frame
  unsigned long long last_call = 0;  // orbx_extractor::call_id of the call that filled the lane
};

struct orbx_extractor {
  int device = 0;
  int nfeatures = 0, nlevels = 0, ini_th = 0, min_th = 0, max_batch = 1;
  float scale_factor = 1.2f;
  std::string err;
  bool planned = false;
  orbx::Plan plan;
  int64_t slab_fstride = 0;
  int in_pitch = 0;
  int64_t in_fstride = 0;
  orbx::ResizeTab* d_tab = nullptr;
  int8_t* d_pattern = nullptr;
  OrbxLane lane[kLanes];
  int last_lane = 0;  // lane of the most recent run
  unsigned long long call_id = 0;  // bumped by every public extract call; "frame f of the last call" = the lane whose
                                   // last_call == call_id and last_f0 <= f < last_f0 + last_frames (api_find_frame)
  // profiling: event records around every stage, resolved lazily by orbx_profile_read (no sync inside a run)
  bool profile = false;
  std::vector<cudaEvent_t> prof_events;  // pool, 2 * kStages per recorded run: (start, end) of every stage
  size_t prof_used = 0;                  // events recorded since the last read
  float prof_ms[kStages] = {};
  int prof_launches[kStages] = {};
};

namespace orbx {
// Bumped whenever the library frees or (re)allocates device memory that recorded work may point to (lane buffers, output
// slabs, matcher scratch). A CUDA graph recorded by the matcher (orbm_api.cu: the small stereo calls) is only replayed
// while the generation it was recorded under is still current.
inline std::atomic<unsigned long long>& alloc_generation() {
  static std::atomic<unsigned long long> g{1};
  return g;
}
// internal entry points of orbx_api.cu used by the matcher's fused stereo call
int api_fail(orbx_extractor* ex, int code, const std::string& msg);
int api_ensure_plan(orbx_extractor* ex, int w, int h);
int api_ensure_out(orbx_extractor* ex, int cap);
// Where frame `frame` of the extractor's most recent call lives: lane + index inside the lane. A call of more than
// kLanes * max_batch frames keeps only its last kLanes groups resident; ORBX_E_ARG for a frame that is gone (or never
// existed).
int api_find_frame(const orbx_extractor* ex, int frame, int* lane, int* local);
// ORBX_TRACE=2: when set, api_upload_and_run records this event right after its H2D copy (timeline of the pipelined calls)
extern thread_local cudaEvent_t g_trace_after_h2d;
// every public extract entry point calls this first: frames of earlier calls stop being addressable
void api_begin_call(orbx_extractor* ex);
// H2D of nb frames into lane `ln` and the whole extractor on stream st; results stay in the lane's device outputs
// copy_stream != nullptr: the H2D copy goes to that stream instead (the pipelined stereo calls keep ALL uploads of a call
// on one stream, so that the copy engine serves them in issue order) and st waits for it through the lane's `up` event
int api_upload_and_run(orbx_extractor* ex, int ln, const uint8_t* src, int nb, int width, int height, int stride,
                       int64_t frame_stride, int lap0, int lap1, cudaStream_t st, int f0 = 0,
                       cudaStream_t copy_stream = nullptr);
// D2H of the lane's outputs of nb frames into rows [0, nb) of the caller's arrays (counts go to the pinned h_small)
int api_download(orbx_extractor* ex, int ln, int nb, orbx_kp* kps, uint8_t* desc, int cap, cudaStream_t st);
}  // namespace orbx

#endif
