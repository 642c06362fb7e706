// orbx_kernels.cuh — internal launch interface between the C ABI (orbx_api.cu) and the sm_100a kernels.
#ifndef ORBX_KERNELS_CUH_
#define ORBX_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orbx_types.h"
#include "orbx_plan.h"

namespace orbx {

// Where the images of one batch live in HBM. Level 0 is read in place from the caller's buffer (or from the H2D
// staging slab); levels >= 1 and every blurred level live in slabs owned by the extractor, one frame after another,
// each level at Plan::lv[l].img_off with Plan::lv[l].pitch.
struct FrameSet {
  const uint8_t* lvl0;
  int pitch0;
  int64_t fstride0;
  uint8_t* pyr;
  uint8_t* blur;
  int64_t slab_fstride;  // = Plan::pyr_bytes_per_frame rounded up
};

__device__ __forceinline__ const uint8_t* raw_level(const Plan& P, const FrameSet& fs, int l, int f, int* pitch) {
  if (l == 0) {
    *pitch = fs.pitch0;
    return fs.lvl0 + (int64_t)f * fs.fstride0;
  }
  *pitch = P.lv[l].pitch;
  return fs.pyr + (int64_t)f * fs.slab_fstride + P.lv[l].img_off;
}
__device__ __forceinline__ uint8_t* blur_level(const Plan& P, const FrameSet& fs, int l, int f) {
  return fs.blur + (int64_t)f * fs.slab_fstride + P.lv[l].img_off;
}

// Per-batch working buffers (device). Sizes per frame come from the Plan.
struct WorkSet {
  uint32_t* slots;        // [frames][slots_per_frame]   FAST candidates per cell, packed x|y|score
  int32_t* cell_count;    // [frames][cells_per_frame]
  uint32_t* cand;         // [frames][slots_per_frame]   per-level compacted candidates (quadtree order)
  uint16_t* lab;          // [frames][slots_per_frame]   quadtree labels (levels whose candidates do not fit in smem)
  uint32_t* lvl_kp;       // [frames][kps_per_frame]     selected keypoints per level, packed x|y|score (ROI-16)
  int32_t* lvl_n;         // [frames][nlevels]           count per level
  int32_t* lvl_c;         // [frames][nlevels]           candidate count per level (diagnostics)
  int32_t* lvl_st;        // [frames][nlevels]           keypoints of the level inside the lapping area ("stereo")
  long long* qt_prof;     // [frames][nlevels][16]       cycle accounting of k_quadtree (ORBX_QT_PROF builds), or null
  int32_t* dst;           // [frames][kps_per_frame]     rank among the level's mono keypoints, or 1<<30 | rank among
                          //                             its stereo keypoints
};

struct OutSet {
  orbx_kp* kps;      // [frames][cap]
  uint8_t* desc;     // [frames][cap][32]
  int32_t* n;        // [frames]
  int32_t* mono;     // [frames]
  int32_t* status;   // [frames] 0 ok, ORBX_E_CAPACITY if n > cap
  int cap;
};

struct ResizeTab {  // one entry per destination row / column
  int16_t ofs, c0, c1, pad;
};

void launch_pyramid(const Plan& P, const FrameSet& fs, const ResizeTab* tab, int frames, cudaStream_t st);
void launch_fast(const Plan& P, const FrameSet& fs, const WorkSet& ws, int ini_th, int min_th, int frames,
                 cudaStream_t st);
void launch_quadtree(const Plan& P, const WorkSet& ws, int lap0, int lap1, int frames, cudaStream_t st);
void launch_blur(const Plan& P, const FrameSet& fs, int frames, cudaStream_t st);
// tensor-core form of the blur (k_blur_tc.cu); false = not applicable (launch_blur then takes k_blur7). ORBX_BLUR_TC=0
// switches it off.
bool launch_blur_tc(const Plan& P, const FrameSet& fs, int frames, cudaStream_t st);
bool blur_tc_enabled();
void launch_describe(const Plan& P, const FrameSet& fs, const WorkSet& ws, const OutSet& out, const int8_t* pattern,
                     int frames, cudaStream_t st);
int launch_cvt_gray(const uint8_t* src, int w, int h, int sstride, int64_t sfstride, int channels, int rgb, uint8_t* dst,
                    int dstride, int64_t dfstride, int frames, cudaStream_t st);
int launch_remap_linear(const uint8_t* src, int sw, int sh, int sstride, int64_t sfstride, const float* mapx,
                        const float* mapy, int dw, int dh, uint8_t* dst, int dstride, int64_t dfstride, int frames,
                        cudaStream_t st);
size_t fast_smem_bytes(const Plan& P);
int fast_launch_count(const Plan& P);  // k_fast launches per extract call (1, or 2 when the small levels are split off)
size_t quadtree_smem_bytes(const Plan& P);

}  // namespace orbx

#endif
