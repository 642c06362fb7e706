// orbx_search_dev.cuh — device helpers shared by the guided searches (k_search.cu) and the batched local-map tracking
// search (k_track.cu): descriptor loads / Hamming distance, the cell window of Frame::GetFeaturesInArea
// (src/Frame.cc:777-801), the per-candidate tests of the search loops, and the (distance, position) top-2 record.
#ifndef ORBX_SEARCH_DEV_CUH_
#define ORBX_SEARCH_DEV_CUH_

#include "orbx_match.cuh"

namespace orbx {

__device__ __forceinline__ void load_desc8(const uint8_t* p, uint32_t (&d)[8]) {
  const uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
  d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ int hamming8(const uint32_t (&a)[8], const uint8_t* p) {
  uint32_t b[8];
  load_desc8(p, b);
  int d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d += __popc(a[i] ^ b[i]);
  return d;
}

struct Window {
  int x0, x1, y0, y1;  // inclusive cell range; x1 < x0 = empty
};

// Frame::GetFeaturesInArea cell range (src/Frame.cc:777-801)
__device__ __forceinline__ Window cell_window(const DevFrame& F, float x, float y, float r) {
  Window w;
  const int C = ORBX_GRID_COLS, R = ORBX_GRID_ROWS;
  w.x0 = max(0, (int)floorf(fmul(fsub(fsub(x, F.min_x), r), F.inv_w)));
  w.x1 = min(C - 1, (int)ceilf(fmul(fadd(fsub(x, F.min_x), r), F.inv_w)));
  w.y0 = max(0, (int)floorf(fmul(fsub(fsub(y, F.min_y), r), F.inv_h)));
  w.y1 = min(R - 1, (int)ceilf(fmul(fadd(fsub(y, F.min_y), r), F.inv_h)));
  if (w.x0 >= C || w.x1 < 0 || w.y0 >= R || w.y1 < 0) w.x1 = w.x0 - 1;
  return w;
}

// all per-candidate tests of the search loops that do not depend on earlier assignments
__device__ __forceinline__ bool cand_ok(const DevFrame& F, int idx, float x, float y, float r, int minLevel,
                                        int maxLevel, bool has_ur, float ur) {
  const orbx_kp kp = F.kps[idx];
  const bool check = (minLevel > 0) || (maxLevel >= 0);  // :803
  if (check) {
    if (kp.octave < minLevel) return false;
    if (maxLevel >= 0 && kp.octave > maxLevel) return false;
  }
  if (!(fabsf(fsub(kp.x, x)) < r && fabsf(fsub(kp.y, y)) < r)) return false;  // :823-826
  if (F.occupied[idx]) return false;                                           // ORBmatcher.cc:92-93 (static part)
  if (has_ur && F.u_right[idx] > 0) {                                          // :95-98
    const float er = fabsf(fsub(ur, F.u_right[idx]));
    if (er > r) return false;
  }
  return true;
}

constexpr int kSearchWarps = 4;

struct Top2 {
  int d1, p1, d2, p2;  // lexicographic (dist, position) minimum and runner-up; p = -1 when absent
};
__device__ __forceinline__ void top2_insert(Top2& t, int d, int p) {
  if (p < 0) return;
  if (t.p1 < 0 || d < t.d1 || (d == t.d1 && p < t.p1)) {
    t.d2 = t.d1; t.p2 = t.p1; t.d1 = d; t.p1 = p;
  } else if (t.p2 < 0 || d < t.d2 || (d == t.d2 && p < t.p2)) {
    t.d2 = d; t.p2 = p;
  }
}
__device__ __forceinline__ Top2 top2_warp(Top2 t) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int d1 = __shfl_xor_sync(0xffffffffu, t.d1, o), p1 = __shfl_xor_sync(0xffffffffu, t.p1, o);
    const int d2 = __shfl_xor_sync(0xffffffffu, t.d2, o), p2 = __shfl_xor_sync(0xffffffffu, t.p2, o);
    top2_insert(t, d1, p1);
    top2_insert(t, d2, p2);
  }
  return t;
}

}  // namespace orbx

#endif
