// k_match.cu — Hamming-search kernels behind include/orbm.h.
//   k_knn2 / k_knn2_merge        cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2)                     src/Frame.cc:1293
//   k_desc_dist                  ORBmatcher::DescriptorDistance                                  src/ORBmatcher.cc:1959-1973
//   k_stereo_match / _median     Frame::ComputeStereoMatches                                     src/Frame.cc:921-1084
// All integer / bitwise work: descriptors are held in registers as 8 x u32, distances are __popc(a ^ b); warp-level
// argmin keeps the reference's tie rules (first minimal element in visiting order wins because it compares with <).
#include <stdlib.h>

#include "orbx_match.cuh"

namespace orbx {

__device__ __forceinline__ void load_desc(const uint8_t* p, uint32_t (&d)[8]) {
  const uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
  d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}

__device__ __forceinline__ int hamming(const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  int d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d += __popc(a[i] ^ b[i]);
  return d;
}

// ---------------------------------------------------------------------------------------------------------------
// knn2: thread = one query (descriptor in registers); the train set streams through shared memory in tiles that every
// thread reads at the same address (broadcast). blockIdx.y splits the train set so that small query sets still fill
// the chip; partial (d1, i1, d2, i2) are merged in split order, which keeps "lower trainIdx wins ties".
// ---------------------------------------------------------------------------------------------------------------
constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 256;  // train rows per shared-memory tile (8 KB)

__device__ __forceinline__ void top2_push(int d, int j, int& b1, int& i1, int& b2, int& i2) {
  if (d < b1) {
    b2 = b1; i2 = i1; b1 = d; i1 = j;
  } else if (d < b2) {
    b2 = d; i2 = j;
  }
}

// Distance of two 256-bit descriptors with 4 POPC instead of 8. POPC issues at a quarter of the LOP3 rate on sm_100a
// (the 8-POPC loop ran at 79 % of that pipe: 457 Gpair/s), so the eight XOR words first go through carry-save adders
// (3-input LOP3: sum = a^b^c, carry = majority): ones + 2*twos + 4*fours, with one word counted on its own.
__device__ __forceinline__ int hamming_csa(const uint32_t (&a)[8], const uint4 u, const uint4 v) {
  const uint32_t x0 = a[0] ^ u.x, x1 = a[1] ^ u.y, x2 = a[2] ^ u.z, x3 = a[3] ^ u.w;
  const uint32_t x4 = a[4] ^ v.x, x5 = a[5] ^ v.y, x6 = a[6] ^ v.z, x7 = a[7] ^ v.w;
  const uint32_t onesA = x0 ^ x1 ^ x2, twosA = (x0 & x1) | (x0 & x2) | (x1 & x2);
  const uint32_t onesB = x3 ^ x4 ^ x5, twosB = (x3 & x4) | (x3 & x5) | (x4 & x5);
  const uint32_t ones = onesA ^ onesB ^ x6, twosC = (onesA & onesB) | (onesA & x6) | (onesB & x6);
  const uint32_t twos = twosA ^ twosB ^ twosC, fours = (twosA & twosB) | (twosA & twosC) | (twosB & twosC);
  return __popc(ones) + __popc(x7) + 2 * __popc(twos) + 4 * __popc(fours);
}

// kPacked: (distance << 22 | train row) keys, so that the running best two are three integer min / max per pair and
// "lower trainIdx wins ties" is the key order itself; needs nt < 2^22 (the launcher falls back otherwise).
template <bool kPacked>
__global__ void __launch_bounds__(kKnnThreads)
k_knn2(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt, int rows_per_split,
       int4* __restrict__ partial) {
  __shared__ uint4 tile[kKnnTile * 2];
  const int qi = blockIdx.x * kKnnThreads + threadIdx.x;
  const int t0 = blockIdx.y * rows_per_split;
  const int t1 = min(nt, t0 + rows_per_split);
  uint32_t a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (qi < nq) load_desc(q + (size_t)qi * 32, a);
  int b1 = 0x7fffffff, b2 = 0x7fffffff, i1 = -1, i2 = -1;
  for (int base = t0; base < t1; base += kKnnTile) {
    const int rows = min(kKnnTile, t1 - base);
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(t + (size_t)base * 32);
    for (int k = threadIdx.x; k < rows * 2; k += kKnnThreads) tile[k] = src[k];
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < rows; r++) {
      const int d = hamming_csa(a, tile[2 * r], tile[2 * r + 1]);
      if constexpr (kPacked) {
        const int key = (d << 22) | (base + r);
        b2 = min(b2, max(b1, key));
        b1 = min(b1, key);
      } else {
        top2_push(d, base + r, b1, i1, b2, i2);
      }
    }
  }
  if constexpr (kPacked) {
    i1 = b1 == 0x7fffffff ? -1 : (b1 & 0x3fffff);
    i2 = b2 == 0x7fffffff ? -1 : (b2 & 0x3fffff);
    b1 = b1 == 0x7fffffff ? b1 : (b1 >> 22);
    b2 = b2 == 0x7fffffff ? b2 : (b2 >> 22);
  }
  if (qi < nq) partial[(size_t)blockIdx.y * nq + qi] = make_int4(b1, i1, b2, i2);
}

__global__ void k_knn2_merge(const int4* __restrict__ partial, int nq, int splits, int32_t* idx1, int32_t* d1,
                             int32_t* idx2, int32_t* d2) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  int b1 = 0x7fffffff, b2 = 0x7fffffff, i1 = -1, i2 = -1;
  for (int s = 0; s < splits; s++) {
    const int4 p = partial[(size_t)s * nq + qi];
    if (p.y >= 0) top2_push(p.x, p.y, b1, i1, b2, i2);
    if (p.w >= 0) top2_push(p.z, p.w, b1, i1, b2, i2);
  }
  idx1[qi] = i1;
  d1[qi] = i1 < 0 ? -1 : b1;
  idx2[qi] = i2;
  d2[qi] = i2 < 0 ? -1 : b2;
}

int knn2_splits(int nq, int nt) {
  const int qblocks = (nq + kKnnThreads - 1) / kKnnThreads;
  int splits = (4 * 148 + qblocks - 1) / qblocks;  // aim at >= 4 CTAs per SM
  const int max_splits = (nt + kKnnTile - 1) / kKnnTile;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

void launch_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int4* partial, int splits, int32_t* idx1,
                 int32_t* d1, int32_t* idx2, int32_t* d2, cudaStream_t st) {
  if (nq <= 0) return;
  int rows = (nt + splits - 1) / splits;
  rows = (rows + kKnnTile - 1) / kKnnTile * kKnnTile;
  if (rows < kKnnTile) rows = kKnnTile;
  dim3 grid((nq + kKnnThreads - 1) / kKnnThreads, splits);
  if (nt < (1 << 22)) k_knn2<true><<<grid, kKnnThreads, 0, st>>>(q, nq, t, nt, rows, partial);
  else k_knn2<false><<<grid, kKnnThreads, 0, st>>>(q, nq, t, nt, rows, partial);
  k_knn2_merge<<<(nq + 255) / 256, 256, 0, st>>>(partial, nq, splits, idx1, d1, idx2, d2);
}

__global__ void k_desc_dist(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x[8], y[8];
  load_desc(a + (size_t)i * 32, x);
  load_desc(b + (size_t)i * 32, y);
  out[i] = hamming(x, y);
}

void launch_knn2_merge(const int4* partial, int nq, int splits, int32_t* idx1, int32_t* d1, int32_t* idx2,
                       int32_t* d2, cudaStream_t st) {
  k_knn2_merge<<<(nq + 255) / 256, 256, 0, st>>>(partial, nq, splits, idx1, d1, idx2, d2);
}

void launch_desc_dist(const uint8_t* a, const uint8_t* b, int n, int32_t* out, cudaStream_t st) {
  if (n > 0) k_desc_dist<<<(n + 255) / 256, 256, 0, st>>>(a, b, n, out);
}

// ---------------------------------------------------------------------------------------------------------------
// The per-feature part of Frame::ComputeBoW (src/Frame.cc:846-851): DBoW2 TemplatedVocabulary::transform(feature,
// word_id, weight, &nid, levelsup) (TemplatedVocabulary.h:1218-1262). One thread per feature walks the tree: at every
// level the first child with the least Hamming distance (strict <), the node passed at level m_L - levelsup is kept
// for the FeatureVector. The 10 children of a node are consecutive rows of the descriptor table, so a level is 20
// independent 16-byte loads; the whole ORB vocabulary (1.1 M nodes x 32 B) sits in the 126 MB L2.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_bow_transform(const DevVocabulary V, const uint8_t* __restrict__ desc, int n, int levelsup,
                uint32_t* __restrict__ word_id, double* __restrict__ weight, uint32_t* __restrict__ node_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[8];
  load_desc(desc + (size_t)i * 32, a);
  const int nid_level = V.depth - levelsup;
  uint32_t nid = 0, final_id = 0;
  int level = 0;
  int c0 = V.child_offsets[0], c1 = V.child_offsets[1];
  do {
    ++level;
    int best_d = 0x7fffffff;
    for (int c = c0; c < c1; c++) {
      const uint32_t id = V.children[c];
      uint32_t b[8];
      load_desc(V.descriptors + (size_t)id * 32, b);
      const int d = hamming(a, b);
      if (d < best_d) {  // the first child wins ties                                      :1242-1251
        best_d = d;
        final_id = id;
      }
    }
    if (level == nid_level) nid = final_id;
    c0 = V.child_offsets[final_id];
    c1 = V.child_offsets[final_id + 1];
  } while (c1 > c0);
  word_id[i] = V.word_id[final_id];
  weight[i] = V.weight[final_id];
  node_id[i] = nid;
}

void launch_bow_transform(const DevVocabulary& V, const uint8_t* desc, int n, int levelsup, uint32_t* word_id,
                          double* weight, uint32_t* node_id, cudaStream_t st) {
  if (n > 0) k_bow_transform<<<(n + 127) / 128, 128, 0, st>>>(V, desc, n, levelsup, word_id, weight, node_id);
}

// ---------------------------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:407-435), batched over map points: CSR of observed
// descriptors -> per point the index of the descriptor with the least median distance to the others. One warp per
// point. For row i the lanes take the columns j = lane, lane + 32, ...; the element [0.5 * (N - 1)] of the sorted row
// is found without sorting: distances are integers in 0..256, so a 257-bin histogram in shared memory + one warp
// scan gives the k-th smallest. The first row with the least median wins (the reference compares with <).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDistinctWarps = 4;
constexpr int kDistinctBins = 288;  // 257 rounded up to 9 bins per lane

__global__ void __launch_bounds__(kDistinctWarps * 32)
k_distinctive(const uint8_t* __restrict__ desc, const int32_t* __restrict__ offsets, int n_points,
              int32_t* __restrict__ best_out) {
  __shared__ int hist_all[kDistinctWarps][kDistinctBins];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * kDistinctWarps + warp;
  if (p >= n_points) return;
  int* hist = hist_all[warp];
  const int o = offsets[p], N = offsets[p + 1] - o;
  if (N <= 0) {
    if (lane == 0) best_out[p] = -1;
    return;
  }
  const uint8_t* D = desc + (size_t)o * 32;
  const int k = (N - 1) >> 1;  // (size_t)(0.5 * (N - 1))                                :429
  int best_median = 0x7fffffff, best_idx = 0;
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int b = 0; b < kDistinctBins / 32; b++) hist[lane * (kDistinctBins / 32) + b] = 0;
    __syncwarp();
    uint32_t a[8];
    load_desc(D + (size_t)i * 32, a);
    for (int j = lane; j < N; j += 32) {
      uint32_t b[8];
      load_desc(D + (size_t)j * 32, b);
      atomicAdd(&hist[hamming(a, b)], 1);
    }
    __syncwarp();
    // k-th smallest (0-based): the first bin whose inclusive prefix count exceeds k
    int mine = 0;
#pragma unroll
    for (int b = 0; b < kDistinctBins / 32; b++) mine += hist[lane * (kDistinctBins / 32) + b];
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, incl > k);
    const int owner = __ffs(hit) - 1;  // N > k, so some lane reaches it
    int median = 0;
    if (lane == owner) {
      int run = incl - mine;
      for (int b = 0; b < kDistinctBins / 32; b++) {
        run += hist[lane * (kDistinctBins / 32) + b];
        if (run > k) {
          median = lane * (kDistinctBins / 32) + b;
          break;
        }
      }
    }
    median = __shfl_sync(0xffffffffu, median, owner);
    if (median < best_median) {  //                                                       :431-434
      best_median = median;
      best_idx = i;
    }
    __syncwarp();
  }
  if (lane == 0) best_out[p] = best_idx;
}

void launch_distinctive(const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best, cudaStream_t st) {
  if (n_points > 0)
    k_distinctive<<<(n_points + kDistinctWarps - 1) / kDistinctWarps, kDistinctWarps * 32, 0, st>>>(desc, offsets,
                                                                                                   n_points, best);
}

// ---------------------------------------------------------------------------------------------------------------
// ComputeStereoMatches. One CTA per (pair, band of kBandRows image rows):
//  1. the reference's row table (right keypoints listed under every row of y +- 2*scale, :939-949) becomes a per-band
//     list in shared memory: the CTA filters the right keypoints whose row span touches its band (a keypoint spans
//     <= ~16 rows, so it lands in 1-2 bands), and the left keypoints whose row int(vL) lies in the band;
//  2. one warp per left keypoint of the band: exact row / octave / u-range test against the band list, 32 candidates
//     per step, Hamming distance of the survivors, lexicographic (distance, iR) argmin — the reference visits the row's
//     candidates in ascending iR with a strict < (:993), i.e. exactly that minimum, so list order is irrelevant;
//  3. 11x11 SAD over 11 shifts on the raw pyramid level of the left keypoint's octave (:1005-1040): lanes own pixels,
//     11 running sums each, then 11 warp reductions;
//  4. parabola fit in non-fused FP32 (:1045-1052), disparity / depth (:1055-1067).
// The first version let every left keypoint scan all right keypoints (1200 x 1200 tests per pair); the band lists cut
// that ~7x. A second kernel (one CTA per pair) finds the median SAD by a two-level histogram select and removes
// matches >= 1.5 * 1.4 * median (:1072-1083).
// ---------------------------------------------------------------------------------------------------------------
// Warps per CTA: 4 for launches that fill the chip (1024 pairs: 1.21 ms; with 8 warps 1.40 ms), 8 for small ones, where
// the CTA's list building (every thread scans 2 x 1200 keypoints / blockDim) and its ~40 left keypoints per band are
// the latency of the launch (single pair: 68 -> 45 us).
constexpr int kStereoWarpsBig = 4, kStereoWarpsSmall = 8;
constexpr int kBandRows = 16;

struct BandEntry {
  int32_t rows;  // minr | maxr << 16
  float x;
  int32_t idx;   // iR | octave << 16
};

template <int kStereoWarps>
__global__ void __launch_bounds__(kStereoWarps * 32)
k_stereo_match(const StereoArgs A) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ int s_nr, s_nl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int band = blockIdx.x, pair = blockIdx.y;
  const int nL = min(A.n_l ? A.n_l[pair] : A.n_l_host, A.cap), nR = min(A.n_r ? A.n_r[pair] : A.n_r_host, A.cap);
  const orbx_kp* kpsL = A.kps_l + (size_t)pair * A.cap;
  const orbx_kp* kpsR = A.kps_r + (size_t)pair * A.cap;
  const uint8_t* descL = A.desc_l + (size_t)pair * A.cap * 32;
  const uint8_t* descR = A.desc_r + (size_t)pair * A.cap * 32;
  float* u_right = A.u_right + (size_t)pair * A.cap;
  float* depth = A.depth + (size_t)pair * A.cap;
  int32_t* sad = A.sad + (size_t)pair * A.cap;
  const int fL = A.frame0 + pair;
  BandEntry* rlist = reinterpret_cast<BandEntry*>(smem_raw);
  uint16_t* llist = reinterpret_cast<uint16_t*>(smem_raw + (size_t)A.cap * sizeof(BandEntry));
  const int nRows = A.left.h[0];
  const int band_lo = band * kBandRows, band_hi = band_lo + kBandRows - 1;
  const float minD = 0.0f, maxD = fdiv(A.mbf, A.mb);
  if (threadIdx.x == 0) { s_nr = 0; s_nl = 0; }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  // ---- right keypoints whose rows [floor(y - r), ceil(y + r)] touch the band (:939-949) ----
  for (int base = 0; base < nR; base += kStereoWarps * 32) {
    const int iR = base + threadIdx.x;
    bool take = false;
    BandEntry e{};
    if (iR < nR) {
      const orbx_kp kpR = kpsR[iR];
      const int oR = kpR.octave;
      if (!(kpR.y == 0.0f && kpR.x == 0.0f) && oR >= 0 && oR < A.nlevels) {  // (0,0) skip: src/Frame.cc:943
        const float r = fmul(2.0f, A.scale[oR]);
        const int maxr = (int)ceilf(fadd(kpR.y, r));
        const int minr = (int)floorf(fsub(kpR.y, r));
        take = maxr >= band_lo && minr <= band_hi;
        e.rows = (minr & 0xffff) | (maxr << 16);
        e.x = kpR.x;
        e.idx = iR | (oR << 16);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    int pos = 0;
    if (lane == 0 && m) pos = atomicAdd(&s_nr, __popc(m));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (take) rlist[pos + __popc(m & lt)] = e;
  }
  // ---- left keypoints of the band; the ones the reference skips (:966-972) are answered here with -1 ----
  for (int base = 0; base < nL; base += kStereoWarps * 32) {
    const int iL = base + threadIdx.x;
    bool take = false;
    if (iL < nL) {
      const orbx_kp kpL = kpsL[iL];
      const int row = (int)kpL.y;  // vRowIndices[vL]: float -> size_t                     :966
      const bool row_ok = row >= 0 && row < nRows;
      if ((row_ok ? row / kBandRows : 0) == band) {
        const float maxU = fsub(kpL.x, minD);
        take = row_ok && !(maxU < 0) && kpL.octave >= 0 && kpL.octave < A.nlevels;
        if (!take) {
          u_right[iL] = -1.0f;
          depth[iL] = -1.0f;
          sad[iL] = -1;
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    int pos = 0;
    if (lane == 0 && m) pos = atomicAdd(&s_nl, __popc(m));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (take) llist[pos + __popc(m & lt)] = (uint16_t)iL;
  }
  __syncthreads();
  const int nBR = s_nr, nBL = s_nl;

  for (int li = warp; li < nBL; li += kStereoWarps) {
    const int iL = llist[li];
    const orbx_kp kpL = kpsL[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    float out_u = -1.0f, out_d = -1.0f;
    int out_sad = -1;
    const float minU = fsub(uL, maxD), maxU = fsub(uL, minD);
    const int row = (int)vL;
    int bestDist = ORBM_TH_HIGH_I, bestIdxR = 0;
    {
      uint32_t dl[8];
      load_desc(descL + (size_t)iL * 32, dl);
      int bd = 0x7fffffff, bi = 0x7fffffff;
      for (int base = 0; base < nBR; base += 32) {
        const int k = base + lane;
        if (k < nBR) {
          const BandEntry e = rlist[k];
          const int minr = (int)(int16_t)(e.rows & 0xffff), maxr = e.rows >> 16;
          const int oR = e.idx >> 16, iR = e.idx & 0xffff;
          if (row >= minr && row <= maxr && !(oR < levelL - 1 || oR > levelL + 1) && e.x >= minU && e.x <= maxU) {
            uint32_t dr[8];
            load_desc(descR + (size_t)iR * 32, dr);
            const int d = hamming(dl, dr);
            if (d < bd || (d == bd && iR < bi)) { bd = d; bi = iR; }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int od = __shfl_xor_sync(0xffffffffu, bd, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
      }
      if (bd < bestDist) { bestDist = bd; bestIdxR = bi; }
    }
    if (bestDist < (ORBM_TH_HIGH_I + ORBM_TH_LOW_I) / 2) {  // thOrbDist                       :925,1001
      const float uR0 = kpsR[bestIdxR].x;
      const float sfac = A.inv_scale[levelL];
      const float scaleduL = roundf(fmul(kpL.x, sfac)), scaledvL = roundf(fmul(kpL.y, sfac));
      const float scaleduR0 = roundf(fmul(uR0, sfac));
      const int w = 5, L = 5;
      const float iniu = fsub(fadd(scaleduR0, (float)L), (float)w);
      const float endu = fadd(fadd(fadd(scaleduR0, (float)L), (float)w), 1.0f);
      const int colsR = A.right.w[levelL];
      bool inb = !(iniu < 0 || endu >= (float)colsR);
      const int yl = (int)fsub(scaledvL, (float)w), xl = (int)fsub(scaleduL, (float)w);
      const int xr0 = (int)fsub(scaleduR0, (float)w);  // inc = 0
      // the reference's cv::Mat ranges throw outside the level; such keypoints cannot come from the extractor
      inb = inb && yl >= 0 && yl + 2 * w < A.left.h[levelL] && yl + 2 * w < A.right.h[levelL] && xl >= 0 &&
            xl + 2 * w < A.left.w[levelL] && xr0 - L >= 0 && xr0 + L + 2 * w < colsR;
      if (inb) {
        const uint8_t* IL = A.left.base[levelL] + (int64_t)fL * A.left.fstride[levelL];
        const uint8_t* IR = A.right.base[levelL] + (int64_t)fL * A.right.fstride[levelL];
        const int pl = A.left.pitch[levelL], pr = A.right.pitch[levelL];
        int acc[11];
#pragma unroll
        for (int k = 0; k < 11; k++) acc[k] = 0;
        for (int p = lane; p < 121; p += 32) {
          const int yy = p / 11, xx = p - yy * 11;
          const int a = IL[(int64_t)(yl + yy) * pl + xl + xx];
          const uint8_t* rrow = IR + (int64_t)(yl + yy) * pr + xr0 + xx - L;
#pragma unroll
          for (int k = 0; k < 11; k++) acc[k] += abs(a - (int)rrow[k]);
        }
        float dists[11];
        int bestSad = 0x7fffffff, bestinc = 0;
#pragma unroll
        for (int k = 0; k < 11; k++) {
          const int s = __reduce_add_sync(0xffffffffu, acc[k]);
          dists[k] = (float)s;          // cv::norm(IL, IR, NORM_L1) -> float               :1033
          if (s < bestSad) { bestSad = s; bestinc = k - L; }
        }
        if (!(bestinc == -L || bestinc == L)) {
          float dist1 = 0, dist2 = 0, dist3 = 0;
#pragma unroll
          for (int k = 1; k < 10; k++)
            if (k == bestinc + L) { dist1 = dists[k - 1]; dist2 = dists[k]; dist3 = dists[k + 1]; }
          const float deltaR = fdiv(fsub(dist1, dist3), fmul(2.0f, fsub(fadd(dist1, dist3), fmul(2.0f, dist2))));
          if (!(deltaR < -1 || deltaR > 1)) {
            float bestuR = fmul(A.scale[levelL], fadd(fadd(scaleduR0, (float)bestinc), deltaR));
            float disparity = fsub(uL, bestuR);
            if (disparity >= minD && disparity < maxD) {
              if (disparity <= 0) {
                disparity = 0.01f;
                bestuR = (float)dsub((double)uL, 0.01);  // float - double literal          :1061
              }
              out_d = fdiv(A.mbf, disparity);
              out_u = bestuR;
              out_sad = bestSad;
            }
          }
        }
      }
    }
    if (lane == 0) {
      u_right[iL] = out_u;
      depth[iL] = out_d;
      sad[iL] = out_sad;
    }
  }
}

// median of the accepted SADs = value of element size/2 of the sorted (dist, iL) vector (:1072-1074) = the
// (m/2)-th smallest SAD; a SAD is < 2^15 (121 * 255), so a two-level histogram select finds it in O(n).
__global__ void __launch_bounds__(256) k_stereo_median(const StereoArgs A) {
  __shared__ int hist[256];
  __shared__ int s_m, s_bin, s_rank, s_median, s_kept;
  const int pair = blockIdx.x;
  const int nL = min(A.n_l ? A.n_l[pair] : A.n_l_host, A.cap);
  float* u_right = A.u_right + (size_t)pair * A.cap;
  float* depth = A.depth + (size_t)pair * A.cap;
  const int32_t* sad = A.sad + (size_t)pair * A.cap;
  hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_m = 0; s_kept = 0; s_median = -1; }
  __syncthreads();
  int cnt = 0;
  for (int i = threadIdx.x; i < nL; i += 256) {
    const int v = sad[i];
    if (v >= 0) {
      atomicAdd(&hist[(v >> 7) & 255], 1);
      cnt++;
    }
  }
  if (cnt) atomicAdd(&s_m, cnt);
  __syncthreads();
  const int m = s_m;
  if (m == 0) {  // the reference reads vDistIdx[0] of an empty vector here (UB); defined as "no matches"
    if (threadIdx.x == 0) A.n_matched[pair] = 0;
    return;
  }
  if (threadIdx.x == 0) {
    int rank = m / 2, b = 0;
    while (rank >= hist[b]) rank -= hist[b++];
    s_bin = b;
    s_rank = rank;
  }
  __syncthreads();
  const int bin = s_bin;
  __syncthreads();
  hist[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < nL; i += 256) {
    const int v = sad[i];
    if (v >= 0 && ((v >> 7) & 255) == bin) atomicAdd(&hist[v & 127], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int rank = s_rank, b = 0;
    while (rank >= hist[b]) rank -= hist[b++];
    s_median = (bin << 7) | b;
  }
  __syncthreads();
  const float median = (float)s_median;
  const float thDist = fmul(0x1.0cccccp+1f, median);  // 1.5f * 1.4f folded to float, then * median   :1074
  int kept = 0;
  for (int i = threadIdx.x; i < nL; i += 256) {
    const int v = sad[i];
    if (v < 0) continue;
    if ((float)v < thDist) kept++;
    else { u_right[i] = -1.0f; depth[i] = -1.0f; }
  }
  if (kept) atomicAdd(&s_kept, kept);
  __syncthreads();
  if (threadIdx.x == 0) A.n_matched[pair] = s_kept;
}

void launch_stereo(const StereoArgs& A, int n_pairs, int max_rows, cudaStream_t st) {
  if (n_pairs <= 0 || max_rows <= 0) return;
  const int bands = (A.left.h[0] + kBandRows - 1) / kBandRows;
  const size_t smem = (size_t)A.cap * (sizeof(BandEntry) + sizeof(uint16_t));
  dim3 grid(bands, n_pairs);
  // small launch = fewer than two CTAs per SM; ORBX_STEREO_WARPS=4|8 forces a form (parity test of one against the other)
  bool small = (long long)bands * n_pairs < 2 * 148;
  if (const char* e = getenv("ORBX_STEREO_WARPS")) small = e[0] == '8';
  if (small) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(k_stereo_match<kStereoWarpsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_stereo_match<kStereoWarpsSmall><<<grid, kStereoWarpsSmall * 32, smem, st>>>(A);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(k_stereo_match<kStereoWarpsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_stereo_match<kStereoWarpsBig><<<grid, kStereoWarpsBig * 32, smem, st>>>(A);
  }
  k_stereo_median<<<n_pairs, 256, 0, st>>>(A);
}

}  // namespace orbx
