// k_search.cu — the guided Hamming searches of ORBmatcher behind include/orbm.h:
//   SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)   src/ORBmatcher.cc:42-221
//   SearchByProjection(Frame&, const Frame&, th, bMono) / (Frame&, KeyFrame*, set, ...)  :1594-1806, :1808-1918
//   SearchForTriangulation(KeyFrame*, KeyFrame*, ...)                                    :886-1106
// with Frame::GetFeaturesInArea (src/Frame.cc:765-831) on the 64x48 grid as a CSR.
//
// Structure of the projection searches:
//   count / fill   one warp per projected point: lanes own the grid cells of the search window, so the candidate list
//                  comes out in the reference's order (ix outer, iy inner, ascending keypoint index inside a cell) from
//                  a prefix sum; the filling pass also computes the Hamming distances and the unconstrained best /
//                  second best per point.
//   resolve        the reference assigns points greedily in vector order and a keypoint taken by an earlier point is
//                  skipped by later ones (:92-93, :130). One warp replays that order: a point whose two best
//                  candidates are both still free keeps its precomputed result (2 shared-memory reads); otherwise the
//                  warp rescans that point's candidate list against the occupancy map.
#include "orbx_match.cuh"
#include "orbx_search_dev.cuh"

namespace orbx {

// FILL = false: counts[i] = number of candidates of point i.
// FILL = true : candidates written at counts[i] (exclusive offsets) + unconstrained top-2 record per point.
template <bool FILL>
__global__ void __launch_bounds__(kSearchWarps * 32)
k_search_enum(const DevFrame F, const DevQueries Q, int32_t* counts, int32_t* cand_idx, int32_t* cand_dist,
              int4* pre) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kSearchWarps + warp;
  if (i >= Q.m) return;
  int total = 0;
  Top2 best{0, -1, 0, -1};
  if (!Q.active || Q.active[i]) {
    const float x = Q.u[i], y = Q.v[i], r = Q.radius[i];
    const int minL = Q.min_level[i], maxL = Q.max_level[i];
    const bool has_ur = Q.u_right != nullptr && F.u_right != nullptr;
    const float ur = has_ur ? Q.u_right[i] : 0.f;
    const Window w = cell_window(F, x, y, r);
    const int ny = w.y1 - w.y0 + 1;
    const int ncell = w.x1 < w.x0 ? 0 : (w.x1 - w.x0 + 1) * ny;
    uint32_t dq[8];
    int out_base = 0;
    if (FILL) {
      load_desc8(Q.desc + (size_t)i * 32, dq);
      out_base = counts[i];
    }
    for (int cb = 0; cb < ncell; cb += 32) {
      const int c = cb + lane;
      int j0 = 0, j1 = 0;
      if (c < ncell) {
        const int ix = w.x0 + c / ny, iy = w.y0 + c % ny;  // ix outer, iy inner
        const int cell = ix * ORBX_GRID_ROWS + iy;
        j0 = F.cell_offsets[cell];
        j1 = F.cell_offsets[cell + 1];
      }
      int n = 0;
      for (int j = j0; j < j1; j++) n += cand_ok(F, F.cell_items[j], x, y, r, minL, maxL, has_ur, ur) ? 1 : 0;
      int inc = n;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (FILL) {
        int pos = total + inc - n;
        for (int j = j0; j < j1; j++) {
          const int idx = F.cell_items[j];
          if (!cand_ok(F, idx, x, y, r, minL, maxL, has_ur, ur)) continue;
          const int dist = hamming8(dq, F.desc + (size_t)idx * 32);
          cand_idx[out_base + pos] = idx;
          cand_dist[out_base + pos] = dist | (F.kps[idx].octave << 16);
          top2_insert(best, dist, pos);
          pos++;
        }
      }
      total += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  if (!FILL) {
    if (lane == 0) counts[i] = total;
  } else {
    best = top2_warp(best);
    if (lane == 0) pre[i] = make_int4(best.d1, best.p1, best.d2, best.p2);
  }
}

void launch_search_count(const DevFrame& F, const DevQueries& Q, int32_t* counts, cudaStream_t st) {
  if (Q.m <= 0) return;
  k_search_enum<false><<<(Q.m + kSearchWarps - 1) / kSearchWarps, kSearchWarps * 32, 0, st>>>(F, Q, counts, nullptr,
                                                                                                nullptr, nullptr);
}

// exclusive scan of counts[0..m) in place, counts[m] = total, *total_out = total (single CTA; m is ~10^4)
__global__ void __launch_bounds__(1024) k_scan(int32_t* counts, int m, int32_t* total_out) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < m; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < m ? counts[i] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int s = warp_sums[lane];
      int sinc = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, sinc, d);
        if (lane >= d) sinc += t;
      }
      warp_sums[lane] = sinc - s;
    }
    __syncthreads();
    const int carry = carry_s;
    if (i < m) counts[i] = carry + warp_sums[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_sums[warp] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[m] = carry_s;
    *total_out = carry_s;
  }
}

void launch_scan(int32_t* counts, int m, int32_t* total_out, cudaStream_t st) {
  k_scan<<<1, 1024, 0, st>>>(counts, m, total_out);
}

void launch_search_fill(const DevFrame& F, const DevQueries& Q, const SearchScratch& S, cudaStream_t st) {
  if (Q.m <= 0) return;
  k_search_enum<true><<<(Q.m + kSearchWarps - 1) / kSearchWarps, kSearchWarps * 32, 0, st>>>(
      F, Q, S.counts, S.cand_idx, S.cand_dist, S.pre);
}

// ---- ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1920-1955) on bin sizes ----
__device__ void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  ind1 = ind2 = ind3 = -1;
  for (int i = 0; i < L; i++) {
    const int s = histo[i];
    if (s > max1) {
      max3 = max2; max2 = max1; max1 = s;
      ind3 = ind2; ind2 = ind1; ind1 = i;
    } else if (s > max2) {
      max3 = max2; max2 = s;
      ind3 = ind2; ind2 = i;
    } else if (s > max3) {
      max3 = s;
      ind3 = i;
    }
  }
  if ((float)max2 < fmul(0.1f, (float)max1)) {
    ind2 = -1;
    ind3 = -1;
  } else if ((float)max3 < fmul(0.1f, (float)max1)) {
    ind3 = -1;
  }
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {
  const float factor = 1.0f / kHistoLength;
  float rot = fsub(a1, a2);
  if (rot < 0.0f) rot = fadd(rot, 360.0f);
  int bin = (int)roundf(fmul(rot, factor));
  if (bin == kHistoLength) bin = 0;
  return bin;
}

// SearchForInitialization (src/ORBmatcher.cc:618-764): every F2 keypoint keeps the closest F1 keypoint seen so far, a
// later F1 keypoint may only take it with a strictly smaller distance (:685) and then evicts the earlier owner
// (:703-706). The candidate lists are produced in parallel by the same count / scan / fill kernels as the other
// searches; this replay of the order dependence is one warp walking the F1 keypoints in index order (the monocular
// initialiser runs a handful of times per session; ~1000 level-0 keypoints).
__global__ void __launch_bounds__(32)
k_init_resolve(const DevFrame F2, const DevQueries Q, const SearchScratch S, const InitArgs A) {
  __shared__ int histo[32];
  const int lane = threadIdx.x;
  for (int k = lane; k < A.n2; k += 32) {
    A.matches21[k] = -1;
    A.matched_dist[k] = 0x7fffffff;
  }
  for (int i = lane; i < Q.m; i += 32) A.matches12[i] = -1;
  histo[lane] = 0;
  __syncwarp();
  int nevents = 0;
  for (int i1 = 0; i1 < Q.m; i1++) {
    const int off = S.counts[i1], cnt = S.counts[i1 + 1] - off;
    if (cnt == 0) continue;  // octave > 0 (:642) or an empty window (:658)
    Top2 t{0, -1, 0, -1};
    for (int c = lane; c < cnt; c += 32) {
      const int dist = S.cand_dist[off + c] & 0xffff;
      if (A.matched_dist[S.cand_idx[off + c]] <= dist) continue;                       // :685
      top2_insert(t, dist, c);
    }
    t = top2_warp(t);
    if (t.p1 < 0) continue;
    const int bestDist = t.d1;
    const float second = t.p2 >= 0 ? (float)t.d2 : (float)0x7fffffff;                  // bestDist2 = INT_MAX when absent
    if (!(bestDist <= ORBM_TH_LOW_I && (float)bestDist < fmul(second, A.nnratio))) continue;  // :697-700
    const int bestIdx2 = S.cand_idx[off + t.p1];
    if (lane == 0) {
      const int prev = A.matches21[bestIdx2];
      if (prev >= 0) A.matches12[prev] = -1;                                            // :703-706
      A.matches12[i1] = bestIdx2;
      A.matches21[bestIdx2] = i1;
      A.matched_dist[bestIdx2] = bestDist;
      if (A.check_orientation) {
        const int bin = rot_bin(A.kps1[i1].angle, F2.kps[bestIdx2].angle);              // :724-734
        A.events[2 * nevents] = i1;
        A.events[2 * nevents + 1] = bin;
        histo[bin]++;
      }
    }
    nevents++;
    __syncwarp();  // lane 0's stores are visible to the whole warp before the next keypoint reads them
  }
  __syncwarp();
  // every eviction took one match away (:705): recount instead of tracking it per step
  int alive = 0;
  for (int i = lane; i < Q.m; i += 32) alive += A.matches12[i] >= 0;
  alive = __reduce_add_sync(0xffffffffu, alive);
  int nmatches = alive;
  if (A.check_orientation) {
    int ind1, ind2, ind3;
    three_maxima(histo, kHistoLength, ind1, ind2, ind3);
    int removed = 0;
    for (int e = lane; e < nevents; e += 32) {
      const int bin = A.events[2 * e + 1], idx1 = A.events[2 * e];
      if (bin != ind1 && bin != ind2 && bin != ind3 && A.matches12[idx1] >= 0) {        // :748-751
        A.matches12[idx1] = -1;
        removed++;
      }
    }
    removed = __reduce_add_sync(0xffffffffu, removed);
    nmatches -= removed;
  }
  if (lane == 0) *A.nmatches = nmatches;
}

void launch_init_resolve(const DevFrame& F2, const DevQueries& Q, const SearchScratch& S, const InitArgs& A,
                         cudaStream_t st) {
  k_init_resolve<<<1, 32, 0, st>>>(F2, Q, S, A);
}

// The greedy, order-dependent part of the searches (src/ORBmatcher.cc:92-93, :130, :1672-1674): point i may only take a
// keypoint that no EARLIER accepted point with observations has taken. Restated as a fixed point: let T[k] be the
// lowest index of an accepted point with observations whose choice is keypoint k; then keypoint k is closed for point
// i iff it was occupied to begin with or T[k] < i, and point i's decision is a pure function of T. Every round
// evaluates all points in parallel against the previous round's T and rebuilds T from the decisions; a decision that
// only depends on lower, already final, decisions is final, so by induction on the index the iteration reaches the
// serial result — and that is its only fixed point — after at most (longest dependency chain) rounds: 3-8 on the
// test maps, where the one-warp serial replay this replaces spent 2.1 ms on 10 000 points. One CTA; T lives in shared
// memory; a point whose two precomputed best candidates are both still open needs no rescan.
constexpr int kResolveThreads = 1024;

struct Decision {
  int idx;    // chosen keypoint, or -1
  float ang;  // its angle (mode 1 with the rotation check)
};

__device__ __forceinline__ Decision resolve_point(const DevFrame& F, const SearchScratch& S, const ResolveArgs& R, int i,
                                                  const int* T, const uint8_t* occ0, bool rot) {
  Decision out{-1, 0.f};
  const int off = S.counts[i], cnt = S.counts[i + 1] - off;
  if (cnt == 0) return out;
  const int4 p = S.pre[i];
  auto closed = [&](int k) { return occ0[k] || T[k] < i; };
  const int k1 = p.y >= 0 ? S.cand_idx[off + p.y] : -1, k2 = p.w >= 0 ? S.cand_idx[off + p.w] : -1;
  int bestDist, bestDist2, bestIdx, bestLevel, bestLevel2;
  if (!((k1 >= 0 && closed(k1)) || (k2 >= 0 && closed(k2)))) {
    // both precomputed best candidates are still open: the filtered search would find the same two
    if (k1 < 0) return out;
    bestDist = p.x;
    bestIdx = k1;
    bestLevel = S.cand_dist[off + p.y] >> 16;
    bestDist2 = k2 >= 0 ? p.z : 256;
    bestLevel2 = k2 >= 0 ? (S.cand_dist[off + p.w] >> 16) : -1;
  } else {
    Top2 t{0, -1, 0, -1};
    for (int c = 0; c < cnt; c++)
      if (!closed(S.cand_idx[off + c])) top2_insert(t, S.cand_dist[off + c] & 0xffff, c);
    if (t.p1 < 0) return out;  // every candidate already taken
    bestDist = t.d1;
    bestIdx = S.cand_idx[off + t.p1];
    bestLevel = S.cand_dist[off + t.p1] >> 16;
    bestDist2 = t.p2 >= 0 ? t.d2 : 256;
    bestLevel2 = t.p2 >= 0 ? (S.cand_dist[off + t.p2] >> 16) : -1;
  }
  bool accept;
  if (R.mode == 0)
    accept = bestDist <= ORBM_TH_HIGH_I &&
             !(bestLevel == bestLevel2 && (float)bestDist > fmul(R.nnratio, (float)bestDist2));  // :124-129
  else
    accept = bestDist <= R.max_dist;  // :1699 / :1897
  if (!accept) return out;
  out.idx = bestIdx;
  if (rot) out.ang = F.kps[bestIdx].angle;
  return out;
}

__global__ void __launch_bounds__(kResolveThreads)
k_search_resolve(const DevFrame F, const DevQueries Q, const SearchScratch S, const ResolveArgs R) {
  extern __shared__ __align__(16) uint8_t rs_smem[];  // int T[n] | int histo[32] | int flag, count | u8 occ0[n]
  const int n = F.n, M = Q.m, tid = threadIdx.x;
  int* T = reinterpret_cast<int*>(rs_smem);
  int* histo = T + n;
  int* flag = histo + 32;
  int* count = flag + 1;
  uint8_t* occ0 = reinterpret_cast<uint8_t*>(count + 1);
  const bool rot = R.mode == 1 && R.check_orientation;
  for (int k = tid; k < n; k += kResolveThreads) {
    T[k] = 0x7fffffff;
    occ0[k] = F.occupied[k];
    R.assign[k] = -1;
  }
  for (int i = tid; i < M; i += kResolveThreads) R.dec[i] = -1;
  if (tid < 32) histo[tid] = 0;
  if (tid == 0) *count = 0;
  __syncthreads();
  for (int round = 0; round <= M; round++) {
    if (tid == 0) *flag = 0;
    __syncthreads();
    bool changed = false;
    for (int i = tid; i < M; i += kResolveThreads) {
      const int d = resolve_point(F, S, R, i, T, occ0, false).idx;
      if (d != R.dec[i]) {
        R.dec[i] = d;
        changed = true;
      }
    }
    if (changed) *flag = 1;
    __syncthreads();
    if (*flag == 0) break;  // fixed point: every decision equals the serial one
    for (int k = tid; k < n; k += kResolveThreads) T[k] = 0x7fffffff;
    __syncthreads();
    for (int i = tid; i < M; i += kResolveThreads) {
      const int d = R.dec[i];
      if (d >= 0 && R.has_obs[i]) atomicMin(&T[d], i);  // a point with observations blocks the keypoint :92-93
    }
    __syncthreads();
  }
  // ---- F.mvpMapPoints[bestIdx] = pMP (:130): the LAST accepted point that chose a keypoint stays; nmatches counts
  //      every acceptance; mode 1 also files every acceptance in the rotation histogram (:1757-1768) ----
  int mine = 0;
  for (int i = tid; i < M; i += kResolveThreads) {
    const int d = R.dec[i];
    if (d < 0) continue;
    atomicMax(&R.assign[d], i);
    mine++;
    if (rot) {
      const int bin = rot_bin(R.angle[i], F.kps[d].angle);
      const int e = atomicAdd(count, 1);
      R.events[2 * e] = d;
      R.events[2 * e + 1] = bin;
      atomicAdd(&histo[bin], 1);
    }
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  __shared__ int warp_sum[kResolveThreads / 32];
  if ((tid & 31) == 0) warp_sum[tid >> 5] = mine;
  __syncthreads();
  int nmatches = 0;
  for (int w = 0; w < kResolveThreads / 32; w++) nmatches += warp_sum[w];
  if (rot) {
    int ind1, ind2, ind3;
    three_maxima(histo, kHistoLength, ind1, ind2, ind3);  // every thread computes the same
    const int nevents = *count;
    __syncthreads();
    if (tid == 0) *count = 0;
    __syncthreads();
    int removed = 0;
    for (int e = tid; e < nevents; e += kResolveThreads) {
      const int bin = R.events[2 * e + 1];
      if (bin != ind1 && bin != ind2 && bin != ind3) {
        R.assign[R.events[2 * e]] = -1;  // :1795-1800
        removed++;
      }
    }
    if (removed) atomicAdd(count, removed);
    __syncthreads();
    nmatches -= *count;
  }
  if (tid == 0) *R.nmatches = nmatches;
}

void launch_search_resolve(const DevFrame& F, const DevQueries& Q, const SearchScratch& S, const ResolveArgs& R,
                           cudaStream_t st) {
  const size_t smem = (size_t)F.n * 4 + 34 * 4 + ((size_t)F.n + 15) / 16 * 16;
  cudaFuncSetAttribute(k_search_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(smem > 48 * 1024 ? smem : 48 * 1024));
  k_search_resolve<<<1, kResolveThreads, smem, st>>>(F, Q, S, R);
}

// ---------------------------------------------------------------------------------------------------------------
// The two-camera form of SearchByProjection(Frame&, vector<MapPoint*>) (Nleft != -1, src/ORBmatcher.cc:42-221). Per
// MapPoint, in vector order: the left search, then — unless the left search left the loop body through its ratio-test
// `continue` (:125-126) — the right-camera twin (:148-217). An accepted point is written to its keypoint AND,
// unconditionally, to the keypoint's stereo partner (mvLeftToRightMatch / mvRightToLeftMatch); a slot is closed for a
// later search while its CURRENT occupant has observations (:92-93, :183-185). Because a partner write can replace an
// occupant that had observations by one that has none, a slot can re-open: "closed" is a function of the last writer, not
// of the first, and the fixed-point iteration of k_search_resolve (which keeps one index per keypoint) does not apply.
// This is the exact serial replay instead: ONE warp walks the points, its lanes share each point's candidate list
// (enumerated and scored in parallel by k_search_enum), slot state lives in shared memory. ~1 ms for 10 000 points; the
// fisheye branch is not on the headline path.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_search_resolve_fisheye(const FisheyeResolveArgs A) {
  extern __shared__ __align__(16) uint8_t fs_smem[];  // int who[N] | u8 blocked[N]
  const int lane = threadIdx.x, N = A.n_left + A.n_right, NL = A.n_left;
  int* who = reinterpret_cast<int*>(fs_smem);
  uint8_t* blocked = reinterpret_cast<uint8_t*>(who + N);
  for (int k = lane; k < N; k += 32) {
    who[k] = -1;
    blocked[k] = A.occupied[k];
  }
  __syncwarp();
  int nmatches = 0;
  for (int i = 0; i < A.m; i++) {
    const uint8_t obs = A.has_obs[i];
    bool skip_right = false;
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
      if (side == 0 ? !A.in_left[i] : (!A.in_right[i] || skip_right)) continue;
      const SearchScratch& S = side == 0 ? A.left : A.right;
      const int base = side == 0 ? 0 : NL;
      const int off = S.counts[i], cnt = S.counts[i + 1] - off;
      if (cnt == 0) continue;
      Top2 t{0, -1, 0, -1};
      for (int c = lane; c < cnt; c += 32)
        if (!blocked[base + S.cand_idx[off + c]]) top2_insert(t, S.cand_dist[off + c] & 0xffff, c);
      t = top2_warp(t);
      if (t.p1 < 0) continue;
      const int bestDist = t.d1, bestLocal = S.cand_idx[off + t.p1], bestLevel = S.cand_dist[off + t.p1] >> 16;
      const int bestDist2 = t.p2 >= 0 ? t.d2 : 256, bestLevel2 = t.p2 >= 0 ? (S.cand_dist[off + t.p2] >> 16) : -1;
      if (bestDist > ORBM_TH_HIGH_I) continue;
      if (bestLevel == bestLevel2 && (float)bestDist > fmul(A.nnratio, (float)bestDist2)) {
        skip_right = side == 0;  // the reference `continue`s the MapPoint loop here
        continue;
      }
      const int partner = side == 0 ? A.left_to_right[bestLocal] : A.right_to_left[bestLocal];
      if (lane == 0) {
        who[base + bestLocal] = i;
        blocked[base + bestLocal] = obs;
        if (partner != -1) {
          const int ps = side == 0 ? partner + NL : partner;
          who[ps] = i;
          blocked[ps] = obs;
        }
      }
      nmatches += 1 + (partner != -1);
      __syncwarp();
    }
  }
  __syncwarp();
  for (int k = lane; k < N; k += 32) A.assign[k] = who[k];
  if (lane == 0) *A.nmatches = nmatches;
}

void launch_search_resolve_fisheye(const FisheyeResolveArgs& A, cudaStream_t st) {
  const size_t N = (size_t)A.n_left + A.n_right;
  const size_t smem = N * 4 + (N + 15) / 16 * 16;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_search_resolve_fisheye, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_search_resolve_fisheye<<<1, 32, smem, st>>>(A);
}

// ---------------------------------------------------------------------------------------------------------------
// SearchForTriangulation. Rows of the result are independent (vbMatched2 is never written in the reference), so:
//   k_tri_nodes   node a of kf1 -> position of the same vocabulary node in kf2 (binary search; both lists ascend)
//   k_tri_match   one thread per shared node replays the reference's double loop for its features
//   k_tri_rot     rotation-consistency histogram + count (one CTA)
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_tri_nodes(const TriArgs A) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.k1.n_nodes) return;
  const uint32_t id = A.k1.node_ids[a];
  int lo = 0, hi = A.k2.n_nodes - 1, found = -1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const uint32_t v = A.k2.node_ids[mid];
    if (v == id) { found = mid; break; }
    if (v < id) lo = mid + 1;
    else hi = mid - 1;
  }
  A.node_match[a] = found;
}

__global__ void k_tri_match(const TriArgs A) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.k1.n_nodes) return;
  const int b = A.node_match[a];
  if (b < 0) return;
  const DevKeyFrame &K1 = A.k1, &K2 = A.k2;
  for (int p1 = K1.offsets[a]; p1 < K1.offsets[a + 1]; p1++) {
    const int idx1 = (int)K1.indices[p1];
    if (K1.has_mappoint[idx1]) continue;                                   // :953-956
    const bool bStereo1 = K1.u_right && K1.u_right[idx1] >= 0;             // :958-960
    if (A.only_stereo && !bStereo1) continue;
    const orbx_kp kp1 = K1.kps[idx1];
    uint32_t d1[8];
    load_desc8(K1.desc + (size_t)idx1 * 32, d1);
    int bestDist = ORBM_TH_LOW_I, bestIdx2 = -1;
    for (int p2 = K2.offsets[b]; p2 < K2.offsets[b + 1]; p2++) {
      const int idx2 = (int)K2.indices[p2];
      if (K2.has_mappoint[idx2]) continue;                                 // :977-980
      const bool bStereo2 = K2.u_right && K2.u_right[idx2] >= 0;
      if (A.only_stereo && !bStereo2) continue;
      const int dist = hamming8(d1, K2.desc + (size_t)idx2 * 32);
      if (dist > ORBM_TH_LOW_I || dist > bestDist) continue;               // :988
      const orbx_kp kp2 = K2.kps[idx2];
      if (!bStereo1 && !bStereo2) {                                        // :996-1003 (pinhole pair)
        const float distex = fsub(A.ep_x, kp2.x), distey = fsub(A.ep_y, kp2.y);
        if (fadd(fmul(distex, distex), fmul(distey, distey)) < fmul(100.f, K2.scale_factors[kp2.octave])) continue;
      }
      bool ok = A.coarse != 0;
      if (!ok) {  // Pinhole::epipolarConstrain, src/CameraModels/Pinhole.cpp:136-148
        const float* F = A.F12;
        const float ea = fadd(fadd(fmul(kp1.x, F[0]), fmul(kp1.y, F[3])), F[6]);
        const float eb = fadd(fadd(fmul(kp1.x, F[1]), fmul(kp1.y, F[4])), F[7]);
        const float ec = fadd(fadd(fmul(kp1.x, F[2]), fmul(kp1.y, F[5])), F[8]);
        const float num = fadd(fadd(fmul(ea, kp2.x), fmul(eb, kp2.y)), ec);
        const float den = fadd(fmul(ea, ea), fmul(eb, eb));
        if (den != 0) {
          const float dsqr = fdiv(fmul(num, num), den);
          ok = (double)dsqr < dmul(3.84, (double)K2.level_sigma2[kp2.octave]);
        }
      }
      if (ok) {
        bestIdx2 = idx2;
        bestDist = dist;
      }
    }
    if (bestIdx2 >= 0) A.matches12[idx1] = bestIdx2;
  }
}

// The descriptor part of SearchForTriangulation alone (:973-988), for rigs whose epipolar test is the caller's
// (KannalaBrandt8::epipolarConstrain triangulates; camera models are not rebuilt): per kf1 feature without a MapPoint the
// kf2 features without a MapPoint under the same vocabulary node with distance <= TH_LOW, in scan order. Pass 1 counts
// (off[idx1] = count, 0 for features that take no part), k_scan turns the counts into offsets, pass 2 writes.
template <bool kFill>
__global__ void k_tri_cand(const TriArgs A, int32_t* off, int32_t* cand_idx2, int32_t* cand_dist, int cap) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.k1.n_nodes) return;
  const int b = A.node_match[a];
  if (b < 0) return;
  const DevKeyFrame &K1 = A.k1, &K2 = A.k2;
  for (int p1 = K1.offsets[a]; p1 < K1.offsets[a + 1]; p1++) {
    const int idx1 = (int)K1.indices[p1];
    if (K1.has_mappoint[idx1]) continue;
    uint32_t d1[8];
    load_desc8(K1.desc + (size_t)idx1 * 32, d1);
    int n = 0;
    const int base = kFill ? off[idx1] : 0;
    for (int p2 = K2.offsets[b]; p2 < K2.offsets[b + 1]; p2++) {
      const int idx2 = (int)K2.indices[p2];
      if (K2.has_mappoint[idx2]) continue;
      const int dist = hamming8(d1, K2.desc + (size_t)idx2 * 32);
      if (dist > ORBM_TH_LOW_I) continue;
      if (kFill && base + n < cap) {
        cand_idx2[base + n] = idx2;
        cand_dist[base + n] = dist;
      }
      n++;
    }
    if (!kFill) off[idx1] = n;
  }
}

void launch_triangulation_candidates(const TriArgs& A, int32_t* off, int32_t* total, int32_t* cand_idx2,
                                     int32_t* cand_dist, int cap, bool fill, cudaStream_t st) {
  const int nb = (A.k1.n_nodes + 63) / 64;
  if (!fill) {
    cudaMemsetAsync(off, 0, (size_t)(A.k1.n + 1) * 4, st);
    if (A.k1.n_nodes > 0) {
      k_tri_nodes<<<(A.k1.n_nodes + 127) / 128, 128, 0, st>>>(A);
      k_tri_cand<false><<<nb, 64, 0, st>>>(A, off, nullptr, nullptr, 0);
    }
    launch_scan(off, A.k1.n, total, st);
  } else if (A.k1.n_nodes > 0) {
    k_tri_cand<true><<<nb, 64, 0, st>>>(A, off, cand_idx2, cand_dist, cap);
  }
}

__global__ void __launch_bounds__(256) k_tri_rot(const TriArgs A) {
  __shared__ int histo[32];
  __shared__ int s_count, s_removed;
  if (threadIdx.x < 32) histo[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_count = 0; s_removed = 0; }
  __syncthreads();
  const int n = A.k1.n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int j = A.matches12[i];
    if (j < 0) continue;
    atomicAdd(&s_count, 1);
    if (A.check_orientation) atomicAdd(&histo[rot_bin(A.k1.kps[i].angle, A.k2.kps[j].angle)], 1);
  }
  __syncthreads();
  if (A.check_orientation) {
    int ind1, ind2, ind3;
    three_maxima(histo, kHistoLength, ind1, ind2, ind3);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int j = A.matches12[i];
      if (j < 0) continue;
      const int bin = rot_bin(A.k1.kps[i].angle, A.k2.kps[j].angle);
      if (bin != ind1 && bin != ind2 && bin != ind3) {
        A.matches12[i] = -1;
        atomicAdd(&s_removed, 1);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *A.nmatches = s_count - s_removed;
}

__global__ void k_fill_i32(int32_t* p, int n, int32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

void launch_triangulation(const TriArgs& A, cudaStream_t st) {
  if (A.k1.n > 0) k_fill_i32<<<(A.k1.n + 255) / 256, 256, 0, st>>>(A.matches12, A.k1.n, -1);
  if (A.k1.n_nodes > 0) {
    k_tri_nodes<<<(A.k1.n_nodes + 127) / 128, 128, 0, st>>>(A);
    k_tri_match<<<(A.k1.n_nodes + 63) / 64, 64, 0, st>>>(A);
  }
  k_tri_rot<<<1, 256, 0, st>>>(A);
}

// ---------------------------------------------------------------------------------------------------------------
// The matching loop of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:1194-1257) after
// the caller-side projection: points are independent (the MapPoint / KeyFrame graph surgery that follows stays with
// the caller, in point order). One warp per point: lanes take the cells of KeyFrame::GetFeaturesInArea's window
// (src/KeyFrame.cc:705-749; ix outer, iy inner) and walk their keypoints; "first keypoint with the least distance"
// (:1253, strict <) is the minimum of (distance, cell position in the window, position in the cell).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSearchWarps * 32)
k_fuse_match(const DevFrame F, const DevQueries Q, const float* __restrict__ inv_level_sigma2, int chi2_gate,
             int32_t* __restrict__ best_idx, int32_t* __restrict__ best_dist) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kSearchWarps + warp;
  if (i >= Q.m) return;
  const float x = Q.u[i], y = Q.v[i], r = Q.radius[i];
  const float ur = Q.u_right ? Q.u_right[i] : 0.f;
  const int L = Q.max_level[i];
  const Window w = cell_window(F, x, y, r);
  const int ny = w.y1 - w.y0 + 1;
  const int ncell = w.x1 < w.x0 ? 0 : (w.x1 - w.x0 + 1) * ny;
  uint32_t dq[8];
  load_desc8(Q.desc + (size_t)i * 32, dq);
  unsigned long long best = ~0ull;  // dist << 32 | window cell << 16 | position in the cell
  int bidx = -1;
  for (int cb = 0; cb < ncell; cb += 32) {
    const int c = cb + lane;
    if (c >= ncell) continue;
    const int ix = w.x0 + c / ny, iy = w.y0 + c % ny;
    const int cell = ix * ORBX_GRID_ROWS + iy;
    const int j0 = F.cell_offsets[cell], j1 = F.cell_offsets[cell + 1];
    for (int j = j0; j < j1; j++) {
      const int idx = F.cell_items[j];
      const orbx_kp kp = F.kps[idx];
      if (!(fabsf(fsub(kp.x, x)) < r && fabsf(fsub(kp.y, y)) < r)) continue;            // KeyFrame.cc:739-743
      const int lv = kp.octave;
      if (lv < L - 1 || lv > L) continue;                                              // :1221
      const float ex = fsub(x, kp.x), ey = fsub(y, kp.y);
      const float kr = F.u_right ? F.u_right[idx] : -1.f;
      if (!chi2_gate) {                                                                // the Sim3 form has none
      } else if (kr >= 0) {                                                            // :1223-1233
        const float er = fsub(ur, kr);
        const float e2 = fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(er, er));
        if ((double)fmul(e2, inv_level_sigma2[lv]) > 7.8) continue;
      } else {                                                                         // :1235-1243
        const float e2 = fadd(fmul(ex, ex), fmul(ey, ey));
        if ((double)fmul(e2, inv_level_sigma2[lv]) > 5.99) continue;
      }
      const int dist = hamming8(dq, F.desc + (size_t)idx * 32);
      const unsigned long long key = ((unsigned long long)dist << 32) | ((unsigned)c << 16) | (unsigned)(j - j0);
      if (key < best) {
        best = key;
        bidx = idx;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob < best) {
      best = ob;
      bidx = oi;
    }
  }
  if (lane == 0) {
    best_idx[i] = bidx;
    best_dist[i] = bidx < 0 ? 256 : (int)(best >> 32);
  }
}

void launch_fuse_match(const DevFrame& F, const DevQueries& Q, const float* inv_level_sigma2, int chi2_gate,
                       int32_t* best_idx, int32_t* best_dist, cudaStream_t st) {
  if (Q.m > 0)
    k_fuse_match<<<(Q.m + kSearchWarps - 1) / kSearchWarps, kSearchWarps * 32, 0, st>>>(F, Q, inv_level_sigma2, chi2_gate,
                                                                                        best_idx, best_dist);
}

// ---------------------------------------------------------------------------------------------------------------
// SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:230-404), Nleft == -1. The only order
// dependence — a frame feature that already received a MapPoint is skipped (:280) — stays inside one vocabulary node,
// because DBoW2 files every feature under exactly one node of the FeatureVector (the ABI checks that the frame's
// lists are disjoint). So: k_bow_nodes pairs the node lists (binary search; both ascend), k_bow_match lets ONE WARP
// PER SHARED NODE take the node's KeyFrame features in order while its lanes share the frame features (a thread per
// node was 4x slower than the CPU on 64-node inputs: 23 x 23 dependent gathers per thread), k_bow_rot applies the
// rotation histogram (:371-388) and counts.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_bow_nodes(const BowArgs A) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.kf.n_nodes) return;
  const uint32_t id = A.kf.node_ids[a];
  int lo = 0, hi = A.fr.n_nodes - 1, found = -1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const uint32_t v = A.fr.node_ids[mid];
    if (v == id) { found = mid; break; }
    if (v < id) lo = mid + 1;
    else hi = mid - 1;
  }
  A.node_match[a] = found;
}

// one WARP per shared node: the node's features of the first view are taken in order (the greedy part), the lanes
// share the node's features of the second view
__global__ void __launch_bounds__(kSearchWarps * 32) k_bow_match(const BowArgs A) {
  const int lane = threadIdx.x & 31;
  const int a = blockIdx.x * kSearchWarps + (threadIdx.x >> 5);
  if (a >= A.kf.n_nodes) return;
  const int b = A.node_match[a];
  if (b < 0) return;
  const DevKeyFrame &K = A.kf, &F = A.fr;
  const int f0 = F.offsets[b], nf = F.offsets[b + 1] - f0;
  for (int pK = K.offsets[a]; pK < K.offsets[a + 1]; pK++) {
    const int idxK = (int)K.indices[pK];
    if (!K.has_mappoint[idxK]) continue;                                   // :262-264 / :802-804
    uint32_t dK[8];
    load_desc8(K.desc + (size_t)idxK * 32, dK);
    Top2 t{0, -1, 0, -1}, tr{0, -1, 0, -1};  // tr: the right camera's rows of a two-camera Frame (:300-311)
    for (int c = lane; c < nf; c += 32) {
      const int idxF = (int)F.indices[f0 + c];
      // taken earlier — only ever by this warp: :280 / :821
      if (A.kf_kf ? (A.matched2[idxF] || !F.has_mappoint[idxF]) : (A.matches_f[idxF] >= 0)) continue;
      const int dist = hamming8(dK, F.desc + (size_t)idxF * 32);
      if (A.n_left_f >= 0 && idxF >= A.n_left_f) top2_insert(tr, dist, c);
      else top2_insert(t, dist, c);                                        // (distance, position) = the strict < scan
    }
    t = top2_warp(t);
    if (A.n_left_f >= 0) {
      // two-camera Frame: the left best passes the ratio test; the right best is taken whenever it is <= TH_LOW, but
      // only inside the left one's "bestDist1 <= TH_LOW" (:319, :346-350 with its "|| true")
      tr = top2_warp(tr);
      if (t.p1 < 0 || t.d1 > ORBM_TH_LOW_I) continue;
      const bool left_ok = (float)t.d1 < fmul(A.nnratio, (float)(t.p2 >= 0 ? t.d2 : 256));
      const bool right_ok = tr.p1 >= 0 && tr.d1 <= ORBM_TH_LOW_I;
      if (lane == 0) {
        if (left_ok) A.matches_f[(int)F.indices[f0 + t.p1]] = idxK;
        if (right_ok) A.matches_f[(int)F.indices[f0 + tr.p1]] = idxK;
      }
      if (left_ok || right_ok) __syncwarp();
      continue;
    }
    if (t.p1 < 0) continue;
    const int best1 = t.d1, best2 = t.p2 >= 0 ? t.d2 : 256;
    const bool low = A.kf_kf ? best1 < ORBM_TH_LOW_I : best1 <= ORBM_TH_LOW_I;  // :838 is strict, :319 is not
    if (low && (float)best1 < fmul(A.nnratio, (float)best2)) {
      const int bestF = (int)F.indices[f0 + t.p1];
      if (lane == 0) {
        if (A.kf_kf) {
          A.matches_f[idxK] = bestF;
          A.matched2[bestF] = 1;
        } else {
          A.matches_f[bestF] = idxK;
        }
      }
      __syncwarp();  // visible to every lane before the next feature of this node is matched
    }
  }
}

__global__ void __launch_bounds__(256) k_bow_rot(const BowArgs A) {
  __shared__ int histo[32];
  __shared__ int s_count, s_removed;
  if (threadIdx.x < 32) histo[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_count = 0; s_removed = 0; }
  __syncthreads();
  const int n = A.kf_kf ? A.kf.n : A.fr.n;
  // rot = angle(KeyFrame / KeyFrame-1 feature) - angle(frame / KeyFrame-2 feature)        :338-344 / :845-851
  auto bin_of = [&](int i, int j) {
    return A.kf_kf ? rot_bin(A.kf.kps[i].angle, A.fr.kps[j].angle) : rot_bin(A.kf.kps[j].angle, A.fr.kps[i].angle);
  };
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int j = A.matches_f[i];
    if (j < 0) continue;
    atomicAdd(&s_count, 1);
    if (A.check_orientation) atomicAdd(&histo[bin_of(i, j)], 1);
  }
  __syncthreads();
  if (A.check_orientation) {
    int ind1, ind2, ind3;
    three_maxima(histo, kHistoLength, ind1, ind2, ind3);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int j = A.matches_f[i];
      if (j < 0) continue;
      const int bin = bin_of(i, j);
      if (bin != ind1 && bin != ind2 && bin != ind3) {
        A.matches_f[i] = -1;
        atomicAdd(&s_removed, 1);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *A.nmatches = s_count - s_removed;
}

__global__ void k_fill_minus_one(int32_t* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = -1;
}

void launch_search_by_bow(const BowArgs& A, cudaStream_t st) {
  const int n_out = A.kf_kf ? A.kf.n : A.fr.n;
  if (n_out > 0) k_fill_minus_one<<<(n_out + 255) / 256, 256, 0, st>>>(A.matches_f, n_out);
  if (A.kf_kf && A.fr.n > 0) cudaMemsetAsync(A.matched2, 0, A.fr.n, st);
  if (A.kf.n_nodes > 0 && A.fr.n_nodes > 0) {
    k_bow_nodes<<<(A.kf.n_nodes + 127) / 128, 128, 0, st>>>(A);
    k_bow_match<<<(A.kf.n_nodes + kSearchWarps - 1) / kSearchWarps, kSearchWarps * 32, 0, st>>>(A);
  }
  k_bow_rot<<<1, 256, 0, st>>>(A);
}

// ---------------------------------------------------------------------------------------------------------------
// Frame::AssignFeaturesToGrid + Frame::PosInGrid (src/Frame.cc:520-547, :833-844), Nleft == -1: the 64x48 lookup grid
// as CSR, cell id = col * 48 + row, keypoint indices ascending inside a cell (push_back order of the serial loop).
// One warp per frame: histogram in shared memory, warp scan over the 3072 cells, then the keypoints are placed in
// index order 32 at a time — lanes that fall into the same cell find each other with match.any and take consecutive
// slots, which keeps every cell's list ascending.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int grid_cell(const orbx_kp& kp, float min_x, float min_y, float inv_w, float inv_h) {
  const int px = (int)roundf(fmul(fsub(kp.x, min_x), inv_w));  // round(): half away from zero              :836-837
  const int py = (int)roundf(fmul(fsub(kp.y, min_y), inv_h));
  if (px < 0 || px >= ORBX_GRID_COLS || py < 0 || py >= ORBX_GRID_ROWS) return -1;  //                        :840-841
  return px * ORBX_GRID_ROWS + py;
}

__global__ void __launch_bounds__(32)
k_build_grid(const orbx_kp* __restrict__ kps, const int32_t* __restrict__ n_ptr, int n_fixed, int64_t kp_stride,
             float min_x, float min_y, float inv_w, float inv_h, int32_t* __restrict__ offsets,
             int32_t* __restrict__ items, int64_t item_stride) {
  constexpr int kCells = ORBX_GRID_COLS * ORBX_GRID_ROWS, kPerLane = kCells / 32;
  __shared__ int32_t cur[kCells];
  const int lane = threadIdx.x, f = blockIdx.x;
  const int n = n_ptr ? n_ptr[f] : n_fixed;
  kps += f * kp_stride;
  offsets += (int64_t)f * (kCells + 1);
  items += f * item_stride;
  for (int c = lane; c < kCells; c += 32) cur[c] = 0;
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const int c = grid_cell(kps[i], min_x, min_y, inv_w, inv_h);
    if (c >= 0) atomicAdd(&cur[c], 1);
  }
  __syncwarp();
  // exclusive scan: lane k owns cells [96k, 96k + 96)
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
    offsets[k] = run;
    cur[k] = run;  // becomes the cell's write cursor
    run += cnt;
  }
  if (lane == 31) offsets[kCells] = run;
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const int c = i < n ? grid_cell(kps[i], min_x, min_y, inv_w, inv_h) : -1;
    const unsigned same = __match_any_sync(0xffffffffu, c);
    int pos = 0;
    if (c >= 0) pos = cur[c];
    __syncwarp();
    if (c >= 0) {
      items[pos + __popc(same & lt)] = i;
      if ((same & lt) == 0u) cur[c] = pos + __popc(same);  // the group's first lane moves the cursor
    }
    __syncwarp();
  }
}

void launch_build_grid(const orbx_kp* kps, const int32_t* n_ptr, int n_fixed, int64_t kp_stride, int frames, float min_x,
                       float min_y, float inv_w, float inv_h, int32_t* offsets, int32_t* items, int64_t item_stride,
                       cudaStream_t st) {
  k_build_grid<<<frames, 32, 0, st>>>(kps, n_ptr, n_fixed, kp_stride, min_x, min_y, inv_w, inv_h, offsets, items,
                                      item_stride);
}

}  // namespace orbx
