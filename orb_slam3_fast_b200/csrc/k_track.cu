// k_track.cu — Tracking::SearchLocalPoints (src/Tracking.cc:3249-3330) for a whole batch of frames that never left the
// device: Frame::isInFrustum (src/Frame.cc:632-699) for every point of each frame's local map, Frame::AssignFeaturesToGrid
// (:520-547), and ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
// (src/ORBmatcher.cc:42-221, Nleft == -1) in its serial MapPoint order. BASELINE.json configs[3].
//
//   k_frustum        thread = (frame, point): the projection and the scalar prologue of the search loop (:55-74:
//                    which points take part, r = RadiusByViewingCos(viewCos) * th * scale[level]) -> one float4 + level
//   k_build_grid     (k_search.cu) one warp per frame
//   k_track_enum     warp = (frame, point): lanes own the cells of the window, candidates come out in the reference's
//                    order; ONE pass: the warp counts, takes its slice of the frame's candidate slab with one atomicAdd,
//                    then writes (distance, octave, keypoint) words and the unconstrained top-2
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
    const int mi = A.map_index ? A.map_index[f] : f % A.n_maps;
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

__device__ __forceinline__ DevFrame track_frame(const TrackArgs& A, int f) {
  DevFrame F;
  F.n = A.n[f];
  F.n_levels = A.n_levels;
  F.kps = A.kps + (size_t)f * A.cap;
  F.desc = A.desc + (size_t)f * A.cap * 32;
  F.u_right = A.u_right ? A.u_right + (size_t)f * A.cap : nullptr;
  F.occupied = A.occupied ? A.occupied + (size_t)f * A.cap : nullptr;
  F.cell_offsets = A.grid_offsets + (size_t)f * (ORBX_GRID_COLS * ORBX_GRID_ROWS + 1);
  F.cell_items = A.grid_items + (size_t)f * A.cap;
  F.min_x = A.min_x;
  F.min_y = A.min_y;
  F.inv_w = A.inv_w;
  F.inv_h = A.inv_h;
  F.scale_factors = nullptr;
  return F;
}

// cand_ok of orbx_search_dev.cuh with an optional occupancy array
__device__ __forceinline__ bool track_cand_ok(const DevFrame& F, int idx, float x, float y, float r, int minLevel,
                                              int maxLevel, bool has_ur, float ur, int* octave) {
  const orbx_kp kp = F.kps[idx];
  *octave = kp.octave;
  if (kp.octave < minLevel || kp.octave > maxLevel) return false;               // Frame.cc:803-817 (maxLevel >= 0 here)
  if (!(fabsf(fsub(kp.x, x)) < r && fabsf(fsub(kp.y, y)) < r)) return false;    // :823-826
  if (F.occupied && F.occupied[idx]) return false;                              // ORBmatcher.cc:92-93 (static part)
  if (has_ur && F.u_right[idx] > 0) {                                           // :95-98
    const float er = fabsf(fsub(ur, F.u_right[idx]));
    if (er > r) return false;
  }
  return true;
}

constexpr int kTrackWarps = 8;

__global__ void __launch_bounds__(kTrackWarps * 32) k_track_enum(const TrackArgs A) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.y;
  const int i = blockIdx.x * kTrackWarps + warp;
  if (i >= A.m) return;
  const size_t o = (size_t)f * A.m + i;
  const int level = A.q_level[o];
  int2 seg = make_int2(0, 0);
  Top2 best{0, -1, 0, -1};
  if (level >= 0) {
    const DevFrame F = track_frame(A, f);
    const float4 q = A.q[o];
    const float x = q.x, y = q.y, ur = q.z, r = q.w;
    // GetFeaturesInArea(x, y, r, level - 1, level): with minLevel = -1 (level 0) the octave test reduces to
    // octave <= maxLevel, which octave >= -1 does not change
    const int minL = level - 1, maxL = level;
    const bool has_ur = F.u_right != nullptr;
    const Window w = cell_window(F, x, y, r);
    const int ny = w.y1 - w.y0 + 1;
    const int ncell = w.x1 < w.x0 ? 0 : (w.x1 - w.x0 + 1) * ny;
    int total = 0, oct;
    for (int cb = 0; cb < ncell; cb += 32) {
      const int c = cb + lane;
      int n = 0;
      if (c < ncell) {
        const int cell = (w.x0 + c / ny) * ORBX_GRID_ROWS + w.y0 + c % ny;  // ix outer, iy inner
        const int j0 = F.cell_offsets[cell], j1 = F.cell_offsets[cell + 1];
        for (int j = j0; j < j1; j++) n += track_cand_ok(F, F.cell_items[j], x, y, r, minL, maxL, has_ur, ur, &oct) ? 1 : 0;
      }
      total += __reduce_add_sync(0xffffffffu, n);
    }
    if (total > 0) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&A.cand_total[f], total);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + total > A.cand_cap) {
        if (lane == 0) A.status[f] = ORBX_E_CAPACITY_I;  // reported, never silently truncated
      } else {
        uint32_t dq[8];
        load_desc8(A.mdesc + ((size_t)(A.map_index ? A.map_index[f] : f % A.n_maps) * A.m + i) * 32, dq);
        uint32_t* out = A.cand + (size_t)f * A.cand_cap + base;
        int run = 0;
        for (int cb = 0; cb < ncell; cb += 32) {
          const int c = cb + lane;
          int j0 = 0, j1 = 0;
          if (c < ncell) {
            const int cell = (w.x0 + c / ny) * ORBX_GRID_ROWS + w.y0 + c % ny;
            j0 = F.cell_offsets[cell];
            j1 = F.cell_offsets[cell + 1];
          }
          int n = 0;
          for (int j = j0; j < j1; j++) n += track_cand_ok(F, F.cell_items[j], x, y, r, minL, maxL, has_ur, ur, &oct) ? 1 : 0;
          int inc = n;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
          }
          int pos = run + inc - n;
          for (int j = j0; j < j1; j++) {
            const int idx = F.cell_items[j];
            if (!track_cand_ok(F, idx, x, y, r, minL, maxL, has_ur, ur, &oct)) continue;
            const int dist = hamming8(dq, F.desc + (size_t)idx * 32);
            out[pos] = ((uint32_t)dist << 20) | ((uint32_t)oct << 16) | (uint32_t)idx;
            top2_insert(best, dist, pos);
            pos++;
          }
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        seg = make_int2(base, total);
      }
    }
  }
  best = top2_warp(best);
  if (lane == 0) {
    A.seg[o] = seg;
    A.pre[o] = make_int4(best.d1, best.p1, best.d2, best.p2);
  }
}

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
  const uint8_t* has_obs = A.has_obs + (size_t)(A.map_index ? A.map_index[f] : f % A.n_maps) * M;
  for (int k = tid; k < n; k += kTrackResolveThreads) {
    T[k] = 0x7fffffff;
    occ0[k] = A.occupied ? A.occupied[(size_t)f * A.cap + k] : 0;
    assign[k] = -1;
  }
  for (int i = tid; i < M; i += kTrackResolveThreads) dec[i] = -1;
  __syncthreads();
  for (int round = 0; round <= M; round++) {
    if (tid == 0) flag = 0;
    __syncthreads();
    bool changed = false;
    for (int i = tid; i < M; i += kTrackResolveThreads) {
      const int2 sg = seg[i];
      if (sg.y == 0) continue;
      const int d = track_resolve_point(A, cand, sg, pre[i], i, T, occ0);
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
    for (int i = tid; i < M; i += kTrackResolveThreads) {
      const int d = dec[i];
      if (d >= 0 && has_obs[i]) atomicMin(&T[d], i);
    }
    __syncthreads();
  }
  // F.mvpMapPoints[bestIdx] = pMP (:130): the last accepted point that chose a keypoint stays; every acceptance counts
  int mine = 0;
  for (int i = tid; i < M; i += kTrackResolveThreads) {
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
  launch_build_grid(A.kps, A.n, 0, A.cap, A.n_frames, A.min_x, A.min_y, A.inv_w, A.inv_h, A.grid_offsets, A.grid_items,
                    A.cap, st);
  cudaMemsetAsync(A.cand_total, 0, (size_t)A.n_frames * 4, st);
  cudaMemsetAsync(A.status, 0, (size_t)A.n_frames * 4, st);
  if (A.m > 0) k_track_enum<<<dim3((A.m + kTrackWarps - 1) / kTrackWarps, A.n_frames), kTrackWarps * 32, 0, st>>>(A);
  const size_t smem = track_resolve_smem(A.cap);
  if (smem > 48 * 1024)  // per device: set on every launch that needs it
    cudaFuncSetAttribute(k_track_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_track_resolve<<<A.n_frames, kTrackResolveThreads, smem, st>>>(A);
}

}  // namespace orbx
