// k_track.cu — Tracking::SearchLocalPoints (src/Tracking.cc:3249-3330) for a whole batch of frames that never left the
// device: Frame::isInFrustum (src/Frame.cc:632-699) for every point of each frame's local map, Frame::AssignFeaturesToGrid
// (:520-547), and ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
// (src/ORBmatcher.cc:42-221, Nleft == -1) in its serial MapPoint order. BASELINE.json configs[3].
//
//   k_frustum        thread = (frame, point): the projection and the scalar prologue of the search loop (:55-74:
//                    which points take part, r = RadiusByViewingCos(viewCos) * th * scale[level]) -> one float4 + level
//   k_track_grid     one warp per frame: the 64x48 grid as u16 offsets + 16-byte search records in cell order
//   k_track_enum     thread = (frame, point) over a frame staged in shared memory (grid, records, descriptors): the
//                    warp takes its points' slices of the frame's candidate slab (sized by the records of each window's
//                    cells) with one atomicAdd, then every thread walks its window once: filter, (distance, octave,
//                    keypoint) words in the reference's candidate order, unconstrained top-2
//                    (a warp per point spent 457 warp instructions per point on mostly idle lanes: 2.83 ms per 256
//                    frames x 10 000 points)
//   k_track_resolve  CTA = frame: the greedy order dependence (:92-93, :130) as the parallel fixed-point iteration of
//                    k_search_resolve, T[] in shared memory
#include "orbx_match.cuh"
#include "orbx_search_dev.cuh"

namespace orbx {

constexpr int ORBX_E_CAPACITY_I = -2;  // ORBX_E_CAPACITY (include/orbx.h)

__global__ void __launch_bounds__(256) k_frustum(const TrackArgs A) {
  const int f = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float fr[26];
  if (threadIdx.x < 26) fr[threadIdx.x] = reinterpret_cast<const float*>(A.frustums + f)[threadIdx.x];
  __shared__ int s_in_view;
  if (threadIdx.x == 0) s_in_view = 0;
  __syncthreads();
  bool counted = false;
  if (i < A.m) {
    const int mi = A.map_index ? A.map_index[f] : (A.map_f0 + f) % A.n_maps;
    const size_t g = (size_t)mi * A.m + i, o = (size_t)f * A.m + i;
    float4 q = make_float4(-1.f, -1.f, 0.f, 0.f);
    int lvl = -1;
    if (A.skip && A.skip[g]) {  // Tracking.cc:3289: not projected at all
      if (A.o_in_view) A.o_in_view[o] = 0;
    } else {
      const int n_levels = reinterpret_cast<const int32_t*>(A.frustums + f)[25];
      const float pos[3] = {A.pos[3 * g], A.pos[3 * g + 1], A.pos[3 * g + 2]};
      const float nrm[3] = {A.normal[3 * g], A.normal[3 * g + 1], A.normal[3 * g + 2]};
      const FrustumOut r = is_in_frustum(fr, n_levels, pos, nrm, A.min_dist[g], A.max_dist[g], A.viewing_cos_limit);
      if (A.o_in_view) A.o_in_view[o] = r.in_view ? 1 : 0;
      if (A.o_proj_x) A.o_proj_x[o] = r.proj_x;
      if (A.o_proj_y) A.o_proj_y[o] = r.proj_y;
      if (r.in_view) {
        counted = true;
        if (A.o_proj_xr) A.o_proj_xr[o] = r.proj_xr;
        if (A.o_view_cos) A.o_view_cos[o] = r.view_cos;
        if (A.o_depth) A.o_depth[o] = r.depth;
        if (A.o_level) A.o_level[o] = r.level;
        // the scalar prologue of the search loop (src/ORBmatcher.cc:55-74, :223-228)
        const bool far = A.far_points && r.depth > A.th_far;
        if (!far && r.level < A.n_levels) {
          float rad = ((double)r.view_cos > 0.998) ? 2.5f : 4.0f;
          if ((double)A.th != 1.0) rad = fmul(rad, A.th);
          q = make_float4(r.proj_x, r.proj_y, r.proj_xr, fmul(rad, A.scale_factors[r.level]));
          lvl = r.level;
        }
      }
    }
    A.q[o] = q;
    A.q_level[o] = lvl;
  }
  const unsigned b = __ballot_sync(0xffffffffu, counted);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_in_view, __popc(b));
  __syncthreads();
  if (threadIdx.x == 0 && s_in_view) atomicAdd(&A.n_in_view[f], s_in_view);
}

void launch_frustum_batch(const TrackArgs& A, cudaStream_t st) {
  cudaMemsetAsync(A.n_in_view, 0, (size_t)A.n_frames * 4, st);
  if (A.m <= 0) return;
  k_frustum<<<dim3((A.m + 255) / 256, A.n_frames), 256, 0, st>>>(A);
}

// ---- Frame::AssignFeaturesToGrid (src/Frame.cc:520-547, PosInGrid :833-844) in the form the enumeration wants: u16 cell
// offsets and, IN CELL ORDER, one 16-byte search record per keypoint {x, y, u_right, keypoint | octave << 16 |
// occupied << 31}, so that walking a cell is a run of consecutive LDS.128. One warp per frame (as k_build_grid).
constexpr int kCells = ORBX_GRID_COLS * ORBX_GRID_ROWS;
constexpr int kOff16 = kTrackOff16;  // u16 offsets per frame: 64 * 48 + 1 used, padded so that a frame's array is 8-byte aligned

// kStage: the frame's keypoints are first copied to shared memory (independent loads, all in flight at once): the two
// passes over them below otherwise pay one dependent global round trip per 32 keypoints (2 x 38 for a 1200-keypoint
// frame: 43 us for a single frame, the longest kernel of the single-pair tracking search).
template <bool kStage>
__global__ void __launch_bounds__(32) k_track_grid(const TrackArgs A) {
  constexpr int kPerLane = kCells / 32;
  __shared__ int32_t cur[kCells];
  extern __shared__ __align__(16) uint8_t tg_smem[];  // kStage: float2 xy[cap] | uint32 meta[cap] | float ur[cap]
  const int lane = threadIdx.x, f = blockIdx.x;
  const int n = min(A.n[f], A.cap);
  const orbx_kp* kps = A.kps + (size_t)f * A.cap;
  const float* ur = A.u_right ? A.u_right + (size_t)f * A.cap : nullptr;
  const uint8_t* occ = A.occupied ? A.occupied + (size_t)f * A.cap : nullptr;
  uint16_t* off16 = A.grid_off16 + (size_t)f * kOff16;
  uint4* rec = A.grid_rec + (size_t)f * A.cap;
  float2* s_xy = reinterpret_cast<float2*>(tg_smem);
  uint32_t* s_meta = reinterpret_cast<uint32_t*>(tg_smem + (size_t)A.cap * 8);
  float* s_ur = reinterpret_cast<float*>(tg_smem + (size_t)A.cap * 12);
  auto cell_xy = [&](float x, float y) {
    const int px = (int)roundf(fmul(fsub(x, A.min_x), A.inv_w));  // round(): half away from zero        :836-837
    const int py = (int)roundf(fmul(fsub(y, A.min_y), A.inv_h));
    if (px < 0 || px >= ORBX_GRID_COLS || py < 0 || py >= ORBX_GRID_ROWS) return -1;  //                  :840-841
    return px * ORBX_GRID_ROWS + py;
  };
  for (int c = lane; c < kCells; c += 32) cur[c] = 0;
  if (kStage) {
#pragma unroll 4
    for (int i = lane; i < n; i += 32) {
      const orbx_kp kp = kps[i];
      s_xy[i] = make_float2(kp.x, kp.y);
      s_meta[i] = (uint32_t)i | ((uint32_t)kp.octave << 16) | ((occ && occ[i]) ? 0x80000000u : 0u);
      s_ur[i] = ur ? ur[i] : -1.f;
    }
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const int c = kStage ? cell_xy(s_xy[i].x, s_xy[i].y) : cell_xy(kps[i].x, kps[i].y);
    if (c >= 0) atomicAdd(&cur[c], 1);
  }
  __syncwarp();
  int sum = 0;
  for (int c = 0; c < kPerLane; c++) sum += cur[lane * kPerLane + c];
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  int run = incl - sum;
  for (int c = 0; c < kPerLane; c++) {
    const int k = lane * kPerLane + c, cnt = cur[k];
    off16[k] = (uint16_t)run;
    cur[k] = run;
    run += cnt;
  }
  if (lane == 31) off16[kCells] = (uint16_t)run;
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    float x = 0, y = 0, u = -1.f;
    uint32_t meta = 0;
    int c = -1;
    if (i < n) {
      if (kStage) {
        x = s_xy[i].x; y = s_xy[i].y; meta = s_meta[i]; u = s_ur[i];
      } else {
        const orbx_kp kp = kps[i];
        x = kp.x; y = kp.y;
        meta = (uint32_t)i | ((uint32_t)kp.octave << 16) | ((occ && occ[i]) ? 0x80000000u : 0u);
        u = ur ? ur[i] : -1.f;
      }
      c = cell_xy(x, y);
    }
    const unsigned same = __match_any_sync(0xffffffffu, c);
    int pos = 0;
    if (c >= 0) pos = cur[c];
    __syncwarp();
    if (c >= 0) {
      rec[pos + __popc(same & lt)] = make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(u), meta);
      if ((same & lt) == 0u) cur[c] = pos + __popc(same);
    }
    __syncwarp();
  }
}

// ---- enumeration: thread = (frame, point). The CTA stages its frame's grid (u16 offsets, records in cell order) and
// descriptors into shared memory once and then serves kEnumPerThread points per thread from there; a thread walks its
// point's window cell by cell (ix outer, iy inner, ascending keypoint index inside a cell = the reference's candidate
// order, src/Frame.cc:803-829) ONCE: the size of its slice of the frame's candidate slab is the number of records in the
// window's cells (known from the offsets; the warp takes the slices of its 32 points with ONE atomicAdd), so the walk
// filters, writes (distance, octave, keypoint) words and keeps the unconstrained top-2 in the same pass. The slab
// capacity (orbx_track_params::cand_per_frame) therefore counts window records, not accepted candidates.
constexpr int kEnumThreads = 256, kEnumPerThread = 8;

__device__ __forceinline__ bool rec_ok(const uint4 r, float x, float y, float rad, int minL, int maxL, bool has_ur,
                                       float ur) {
  const int oct = (int)((r.w >> 16) & 0x7fffu);
  if (oct < minL || oct > maxL) return false;                                         // Frame.cc:803-817
  const float kx = __uint_as_float(r.x), ky = __uint_as_float(r.y);
  if (!(fabsf(fsub(kx, x)) < rad && fabsf(fsub(ky, y)) < rad)) return false;          // :823-826
  if (r.w & 0x80000000u) return false;                                                // ORBmatcher.cc:92-93 (static part)
  const float kur = __uint_as_float(r.z);
  if (has_ur && kur > 0) {                                                            // :95-98
    if (fabsf(fsub(ur, kur)) > rad) return false;
  }
  return true;
}

__global__ void __launch_bounds__(kEnumThreads) k_track_enum(const TrackArgs A) {
  extern __shared__ __align__(16) uint8_t en_smem[];  // uint4 rec[cap] | uint4 desc[2 * cap] | u16 off[kOff16]
  const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int n = min(A.n[f], A.cap);
  uint4* s_rec = reinterpret_cast<uint4*>(en_smem);
  uint4* s_desc = s_rec + A.cap;
  uint16_t* s_off = reinterpret_cast<uint16_t*>(s_desc + 2 * (size_t)A.cap);
  {
    const uint4* g_rec = A.grid_rec + (size_t)f * A.cap;
    const uint4* g_desc = reinterpret_cast<const uint4*>(A.desc + (size_t)f * A.cap * 32);
    const uint32_t* g_off = reinterpret_cast<const uint32_t*>(A.grid_off16 + (size_t)f * kOff16);
    for (int i = tid; i < n; i += kEnumThreads) s_rec[i] = g_rec[i];
    for (int i = tid; i < 2 * n; i += kEnumThreads) s_desc[i] = g_desc[i];
    for (int i = tid; i < kOff16 / 2; i += kEnumThreads) reinterpret_cast<uint32_t*>(s_off)[i] = g_off[i];
  }
  __syncthreads();
  const bool has_ur = A.u_right != nullptr;
  const size_t map_base = (size_t)(A.map_index ? A.map_index[f] : (A.map_f0 + f) % A.n_maps) * A.m;
  uint32_t* slab = A.cand + (size_t)f * A.cand_cap;
  const int i0 = blockIdx.x * (kEnumThreads * kEnumPerThread);
  // The CTA's points are visited in the order of their predicted level (a counting sort of <= 2048 local indices): a
  // level-7 window covers 13x the area of a level-0 one, and a warp whose lanes hold random levels runs every lane at
  // the pace of its largest window. Points that are not searched (not in view, far, skipped) go last: whole warps
  // then have nothing to do. The order of the points is free — every point owns its slice of the slab.
  __shared__ int s_cnt[kMaxLevels + 1], s_base[kMaxLevels + 1];
  __shared__ uint16_t s_order[kEnumThreads * kEnumPerThread];
  if (tid <= kMaxLevels) s_cnt[tid] = 0;
  __syncthreads();
  {
    int key[kEnumPerThread], rank[kEnumPerThread];
#pragma unroll
    for (int k = 0; k < kEnumPerThread; k++) {
      const int i = i0 + k * kEnumThreads + tid;
      const int lv = i < A.m ? A.q_level[(size_t)f * A.m + i] : -1;
      key[k] = lv < 0 ? kMaxLevels : lv;
      rank[k] = atomicAdd(&s_cnt[key[k]], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int b = 0; b <= kMaxLevels; b++) {
        s_base[b] = run;
        run += s_cnt[b];
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kEnumPerThread; k++) s_order[s_base[key[k]] + rank[k]] = (uint16_t)(k * kEnumThreads + tid);
  }
  __syncthreads();
  const int n_active = s_base[kMaxLevels];  // searched points of this CTA
#pragma unroll 1
  for (int k = 0; k < kEnumPerThread; k++) {
    const int slot = k * kEnumThreads + tid;
    if (k * kEnumThreads >= n_active) break;  // CTA-uniform: nothing but unsearched points from here on
    const int i = i0 + (int)s_order[slot];
    const bool valid = slot < n_active;
    const size_t o = (size_t)f * A.m + (valid ? i : 0);
    const int level = valid ? A.q_level[o] : -1;
    float x = 0, y = 0, ur = 0, rad = 0;
    int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
    const int minL = level - 1, maxL = level;  // GetFeaturesInArea(x, y, r, level - 1, level)
    // Upper bound of the point's candidate count = the records of its window's cells, from the offsets alone (no record
    // is read): the point's slice of the slab is sized by it, so ONE walk over the records serves both the filter and
    // the distances (a counting pass first, as round 2's first version did, read every record twice)
    int wnd = 0;
    if (level >= 0) {
      const float4 q = A.q[o];
      x = q.x; y = q.y; ur = q.z; rad = q.w;
      // Frame::GetFeaturesInArea's cell range (src/Frame.cc:777-801)
      x0 = max(0, (int)floorf(fmul(fsub(fsub(x, A.min_x), rad), A.inv_w)));
      x1 = min(ORBX_GRID_COLS - 1, (int)ceilf(fmul(fadd(fsub(x, A.min_x), rad), A.inv_w)));
      y0 = max(0, (int)floorf(fmul(fsub(fsub(y, A.min_y), rad), A.inv_h)));
      y1 = min(ORBX_GRID_ROWS - 1, (int)ceilf(fmul(fadd(fsub(y, A.min_y), rad), A.inv_h)));
      if (x0 >= ORBX_GRID_COLS || x1 < 0 || y0 >= ORBX_GRID_ROWS || y1 < 0) x1 = x0 - 1;
      // the cells (ix, y0..y1) are consecutive: one run of records per column
      for (int ix = x0; ix <= x1; ix++) wnd += s_off[ix * ORBX_GRID_ROWS + y1 + 1] - s_off[ix * ORBX_GRID_ROWS + y0];
    }
    uint32_t dq[8];
    if (wnd > 0) load_desc8(A.mdesc + (map_base + i) * 32, dq);  // requested before the scan below hides its latency
    // the warp takes its slice of the frame's slab
    int inc = wnd;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    const int wtotal = __shfl_sync(0xffffffffu, inc, 31);
    int wbase = 0;
    if (lane == 0 && wtotal > 0) wbase = atomicAdd(&A.cand_total[f], wtotal);
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    const bool overflow = wbase + wtotal > A.cand_cap;
    if (overflow && lane == 0) A.status[f] = ORBX_E_CAPACITY_I;  // reported, never silently truncated
    if (!valid) continue;
    int2 seg = make_int2(0, 0);
    Top2 best{0, -1, 0, -1};
    if (wnd > 0 && !overflow) {
      const int base = wbase + inc - wnd;
      int pos = 0;
      for (int ix = x0; ix <= x1; ix++) {
        const int j0 = s_off[ix * ORBX_GRID_ROWS + y0], j1 = s_off[ix * ORBX_GRID_ROWS + y1 + 1];
        for (int j = j0; j < j1; j++) {
          const uint4 r = s_rec[j];
          if (!rec_ok(r, x, y, rad, minL, maxL, has_ur, ur)) continue;
          const int idx = (int)(r.w & 0xffffu);
          const uint4 da = s_desc[2 * idx], db = s_desc[2 * idx + 1];
          const int dist = __popc(dq[0] ^ da.x) + __popc(dq[1] ^ da.y) + __popc(dq[2] ^ da.z) + __popc(dq[3] ^ da.w) +
                           __popc(dq[4] ^ db.x) + __popc(dq[5] ^ db.y) + __popc(dq[6] ^ db.z) + __popc(dq[7] ^ db.w);
          slab[base + pos] = ((uint32_t)dist << 20) | (r.w & 0x000fffffu);
          top2_insert(best, dist, pos);
          pos++;
        }
      }
      if (pos > 0) seg = make_int2(base, pos);
    }
    A.seg[o] = seg;
    A.pre[o] = make_int4(best.d1, best.p1, best.d2, best.p2);
  }
  // unsearched points: an empty candidate list
  for (int slot = n_active + tid; slot < kEnumThreads * kEnumPerThread; slot += kEnumThreads) {
    const int i = i0 + (int)s_order[slot];
    if (i < A.m) A.seg[(size_t)f * A.m + i] = make_int2(0, 0);
  }
}

size_t track_enum_smem(int cap) { return (size_t)cap * 48 + ((size_t)kOff16 * 2 + 15) / 16 * 16; }

constexpr int kTrackResolveThreads = 1024;

__device__ __forceinline__ int track_resolve_point(const TrackArgs& A, const uint32_t* cand, int2 sg, int4 p, int i,
                                                   const int* T, const uint8_t* occ0) {
  if (sg.y == 0) return -1;
  const uint32_t* c = cand + sg.x;
  auto closed = [&](int k) { return occ0[k] || T[k] < i; };
  const uint32_t w1 = p.y >= 0 ? c[p.y] : 0u, w2 = p.w >= 0 ? c[p.w] : 0u;
  const int k1 = p.y >= 0 ? (int)(w1 & 0xffffu) : -1, k2 = p.w >= 0 ? (int)(w2 & 0xffffu) : -1;
  int bestDist, bestDist2, bestIdx, bestLevel, bestLevel2;
  if (!((k1 >= 0 && closed(k1)) || (k2 >= 0 && closed(k2)))) {
    if (k1 < 0) return -1;
    bestDist = p.x;
    bestIdx = k1;
    bestLevel = (int)((w1 >> 16) & 0xfu);
    bestDist2 = k2 >= 0 ? p.z : 256;
    bestLevel2 = k2 >= 0 ? (int)((w2 >> 16) & 0xfu) : -1;
  } else {
    Top2 t{0, -1, 0, -1};
    for (int e = 0; e < sg.y; e++)
      if (!closed((int)(c[e] & 0xffffu))) top2_insert(t, (int)(c[e] >> 20), e);
    if (t.p1 < 0) return -1;
    bestDist = t.d1;
    bestIdx = (int)(c[t.p1] & 0xffffu);
    bestLevel = (int)((c[t.p1] >> 16) & 0xfu);
    bestDist2 = t.p2 >= 0 ? t.d2 : 256;
    bestLevel2 = t.p2 >= 0 ? (int)((c[t.p2] >> 16) & 0xfu) : -1;
  }
  const bool accept = bestDist <= ORBM_TH_HIGH_I &&
                      !(bestLevel == bestLevel2 && (float)bestDist > fmul(A.nnratio, (float)bestDist2));  // :124-129
  return accept ? bestIdx : -1;
}

__global__ void __launch_bounds__(kTrackResolveThreads) k_track_resolve(const TrackArgs A) {
  extern __shared__ __align__(16) uint8_t tr_smem[];  // int T[cap] | u8 occ0[cap]
  __shared__ int flag, warp_sum[kTrackResolveThreads / 32];
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n = min(A.n[f], A.cap), M = A.m;
  int* T = reinterpret_cast<int*>(tr_smem);
  uint8_t* occ0 = reinterpret_cast<uint8_t*>(T + A.cap);
  int32_t* assign = A.assign + (size_t)f * A.cap;
  int32_t* dec = A.dec + (size_t)f * M;
  const int2* seg = A.seg + (size_t)f * M;
  const int4* pre = A.pre + (size_t)f * M;
  const uint32_t* cand = A.cand + (size_t)f * A.cand_cap;
  const uint8_t* has_obs = A.has_obs + (size_t)(A.map_index ? A.map_index[f] : (A.map_f0 + f) % A.n_maps) * M;
  for (int k = tid; k < n; k += kTrackResolveThreads) {
    T[k] = 0x7fffffff;
    occ0[k] = A.occupied ? A.occupied[(size_t)f * A.cap + k] : 0;
    assign[k] = -1;
  }
  // only points with candidates take part in the iteration: their indices are compacted into the (no longer needed)
  // q_level slice of the frame; the order inside the list is irrelevant, a point's rank in the serial order is its index
  int32_t* act = A.q_level + (size_t)f * M;
  __shared__ int n_act_s;
  if (tid == 0) n_act_s = 0;
  __syncthreads();
  for (int base = 0; base < M; base += kTrackResolveThreads) {
    const int i = base + tid;
    const bool has = i < M && seg[i].y > 0;
    const unsigned bal = __ballot_sync(0xffffffffu, has);
    int wbase = 0;
    if ((tid & 31) == 0 && bal) wbase = atomicAdd(&n_act_s, __popc(bal));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (i < M) dec[i] = -1;
    __syncthreads();  // the reads of seg[] / q_level above precede the writes into act[] (same memory as q_level)
    if (has) act[wbase + __popc(bal & ((1u << (tid & 31)) - 1u))] = i;
  }
  __syncthreads();
  const int n_act = n_act_s;
  for (int round = 0; round <= n_act; round++) {
    if (tid == 0) flag = 0;
    __syncthreads();
    bool changed = false;
    for (int a = tid; a < n_act; a += kTrackResolveThreads) {
      const int i = act[a];
      const int d = track_resolve_point(A, cand, seg[i], pre[i], i, T, occ0);
      if (d != dec[i]) {
        dec[i] = d;
        changed = true;
      }
    }
    if (changed) flag = 1;
    __syncthreads();
    if (flag == 0) break;
    for (int k = tid; k < n; k += kTrackResolveThreads) T[k] = 0x7fffffff;
    __syncthreads();
    for (int a = tid; a < n_act; a += kTrackResolveThreads) {
      const int i = act[a];
      const int d = dec[i];
      if (d >= 0 && has_obs[i]) atomicMin(&T[d], i);
    }
    __syncthreads();
  }
  // F.mvpMapPoints[bestIdx] = pMP (:130): the last accepted point that chose a keypoint stays; every acceptance counts
  int mine = 0;
  for (int a = tid; a < n_act; a += kTrackResolveThreads) {
    const int i = act[a];
    const int d = dec[i];
    if (d < 0) continue;
    atomicMax(&assign[d], i);
    mine++;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((tid & 31) == 0) warp_sum[tid >> 5] = mine;
  __syncthreads();
  if (tid == 0) {
    int nm = 0;
    for (int w = 0; w < kTrackResolveThreads / 32; w++) nm += warp_sum[w];
    // SearchLocalPoints only calls the matcher when nToMatch > 0 (Tracking.cc:3301); with no point in view nm is 0 anyway
    A.nmatches[f] = nm;
  }
}

size_t track_resolve_smem(int cap) { return (size_t)cap * 4 + ((size_t)cap + 15) / 16 * 16; }

void launch_track_search(const TrackArgs& A, cudaStream_t st) {
  cudaMemsetAsync(A.cand_total, 0, (size_t)A.n_frames * 4, st);
  cudaMemsetAsync(A.status, 0, (size_t)A.n_frames * 4, st);
  {
    const size_t gs = (size_t)A.cap * 16;  // staged keypoints: xy, meta, u_right
    if (gs <= 160 * 1024) {
      if (gs > 36 * 1024) cudaFuncSetAttribute(k_track_grid<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs);
      k_track_grid<true><<<A.n_frames, 32, gs, st>>>(A);
    } else {
      k_track_grid<false><<<A.n_frames, 32, 0, st>>>(A);
    }
  }
  if (A.m > 0) {
    const size_t es = track_enum_smem(A.cap);
    if (es > 48 * 1024) cudaFuncSetAttribute(k_track_enum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)es);
    const int per_cta = kEnumThreads * kEnumPerThread;
    k_track_enum<<<dim3((A.m + per_cta - 1) / per_cta, A.n_frames), kEnumThreads, es, st>>>(A);
  }
  const size_t smem = track_resolve_smem(A.cap);
  if (smem > 48 * 1024)  // per device: set on every launch that needs it
    cudaFuncSetAttribute(k_track_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_track_resolve<<<A.n_frames, kTrackResolveThreads, smem, st>>>(A);
}

}  // namespace orbx
