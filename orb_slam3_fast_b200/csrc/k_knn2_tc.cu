// k_knn2_tc.cu — tensor-core form of cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2) (src/Frame.cc:1293; SURVEY.md §8(f)
// rank 4) for large query x train sets; k_knn2 in k_match.cu stays the path for small ones.
//
// A 256-bit descriptor is expanded once to 256 signed bytes of +-1 (32 B -> 256 B per row); then
//   dot(a', b') = 256 - 2 * hamming(a, b)      exactly, in int32,
// i.e. a plain s8 x s8 -> s32 GEMM: tcgen05.mma kind::i8, M = 128 queries x N = 256 train rows x K = 256 per tile,
// operands K-major in 128B-swizzled shared memory written by TMA, the accumulator double-buffered in TMEM (2 x 256
// columns) so that the epilogue of tile i overlaps the MMAs of tile i + 1. The epilogue owns one query row per thread
// (= TMEM lane), reads 32 accumulator columns per tcgen05.ld and keeps the best two as packed
// (distance << 22 | train row) keys: the same keys, and therefore the same "lower trainIdx wins ties" order, as k_knn2.
// Train tiles are visited in ascending row order, so a candidate can only enter the best two if its distance is
// STRICTLY below the current second best: a 32-column chunk is skipped after one 3-input max reduction (half an
// instruction per value) unless its largest accumulator beats that threshold; the exact insertion runs on ~2 ln(nt)
// chunks per row. One CTA per SM (160 KB of shared memory, all 512 TMEM columns): warp 0 = TMA producer, warp 1 = MMA
// issuer, warps 2..5 = epilogue (warp w may touch TMEM lanes 32 (w % 4) .. +31). blockIdx.y splits the train set;
// the per-split (d1, i1, d2, i2) are merged by k_knn2_merge in split order like the POPC kernel's.
//
// Two forms live here. k_knn2_tc (this description) was round 2's first: 1.70 ms for 100k x 100k. k_knn2_tc_ts further
// down (256 queries per CTA, both query tiles in tensor memory, N = 192, four train tiles in flight, two epilogue warps
// per TMEM lane quarter) is the one that runs: 1.50 ms = 6.66 Tpair/s = 3.4 Pop/s of s8 MACs, 76 % of the nominal
// 4.5 Pop/s dense int8 peak (bench.py --config 4, device-resident, expansion and merge included; POPC kernel: 15.0 ms).
// History of the first form: 2.20 ms with one tcgen05.wait::ld per 32-column load, 1.77 ms with four in flight, 1.70 ms
// with the lean issue path; 3.66 ms without the chunk filter; through the library in round 1: 2.78 ms (7 train-set
// splits, see knn2_tc_splits). ORBM_KNN2_TS=0 selects it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "orbx_match.cuh"

namespace orbx {
namespace {
constexpr int kM = 128, kN = 256, kRowBytes = 256, kHalf = 128;  // K = 256 bytes per row = two 128-byte swizzle spans
constexpr int kABytes = kM * kRowBytes, kBBytes = kN * kRowBytes;
constexpr int kThreads = 192;
constexpr int kSmem = kABytes + 2 * kBBytes + 256 + 1024;  // + barriers + alignment slack
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);

__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(b)) : "memory");
}
// Bounded spin (~seconds): see the trap below.
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  for (uint32_t spins = 0;; spins++) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(sptr(b)), "r"(parity) : "memory");
    if (ok) return;
    if (spins > (1u << 24)) {
      __trap();  // a protocol error must surface as a CUDA error on the caller's stream, never as a hung device
    }
  }
}
__device__ __forceinline__ void tma_rows(const CUtensorMap* map, void* dst, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(sptr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(sptr(bar)) : "memory");
}
// K-major operand in SWIZZLE_128B layout: 8-row groups 1024 B apart (SBO = 64 x 16 B), LBO unused (1), version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(const void* p) {
  const uint64_t a = (sptr(p) & 0x3ffff) >> 4;
  return a | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sptr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, int (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 B of bits -> 256 B of +-1 (bit j of byte b -> element 8 b + j; any fixed order works, both sides use the same)
__global__ void k_expand_pm1(const uint8_t* __restrict__ desc, int n, int8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 32) return;
  const uint32_t b = desc[i];
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    lo |= (((b >> j) & 1) ? 0x01u : 0xffu) << (8 * j);
    hi |= (((b >> (4 + j)) & 1) ? 0x01u : 0xffu) << (8 * j);
  }
  reinterpret_cast<uint2*>(out)[i] = make_uint2(lo, hi);
}

// kGroups = 32-column TMEM loads in flight per tcgen05.wait::ld (1 / 2 / 4 measured in tools/ubench/knn2_tc.cu: 2.20 /
// 1.85 / 1.77 ms for 100k x 100k — one wait per load left the epilogue warps idle for the TMEM round trip 8 times a tile)
template <bool kFilter, int kGroups>
__global__ void __launch_bounds__(kThreads, 1)
k_knn2_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t, int nq, int nt,
          int tiles_per_split, int4* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // [2 halves][128 rows][128 B]
  uint8_t* sB = smem + kABytes;             // [2 stages][2 halves][256 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kABytes + 2 * kBBytes);
  uint64_t *a_full = bars, *b_full = bars + 1, *b_empty = bars + 3, *acc_full = bars + 5, *acc_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  // warp index and TMEM base are broadcast with shfl so that ptxas can prove them warp-uniform: the issue paths below
  // run on the whole warp, only the asynchronous instructions sit under elect.sync. (Under `if (lane == 0)` every
  // UTCIMMA / UTMALDG was wrapped in a uniformisation loop — R2UR.BROADCAST + BRA.U.ANY, ~13 instructions per MMA.)
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int q0 = blockIdx.x * kM;
  const int total_tiles = (nt + kN - 1) / kN;
  const int tb = blockIdx.y * tiles_per_split;
  const int ntile = max(0, min(total_tiles, tb + tiles_per_split) - tb);

  if (threadIdx.x == 0) {
    bar_init(a_full, 1);
    for (int s = 0; s < 2; s++) {
      bar_init(b_full + s, 1);
      bar_init(b_empty + s, 1);
      bar_init(acc_full + s, 1);
      bar_init(acc_empty + s, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (elect_one()) {
      bar_expect(a_full, kABytes);
      tma_rows(&map_q, sA, a_full, 0, q0);
      tma_rows(&map_q, sA + kM * kHalf, a_full, kHalf, q0);
    }
    __syncwarp();
    for (int i = 0; i < ntile; i++) {
      const int s = i & 1;
      bar_wait(b_empty + s, ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        bar_expect(b_full + s, kBBytes);
        tma_rows(&map_t, sB + s * kBBytes, b_full + s, 0, (tb + i) * kN);
        tma_rows(&map_t, sB + s * kBBytes + kN * kHalf, b_full + s, kHalf, (tb + i) * kN);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    uint64_t da[8], db[2][8];
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {  // 8 x (K = 32 bytes); 4 steps inside each 128-byte swizzle span
      da[kk] = smem_desc(sA + (kk >> 2) * (kM * kHalf) + (kk & 3) * 32);
      db[0][kk] = smem_desc(sB + (kk >> 2) * (kN * kHalf) + (kk & 3) * 32);
      db[1][kk] = smem_desc(sB + kBBytes + (kk >> 2) * (kN * kHalf) + (kk & 3) * 32);
    }
    bar_wait(a_full, 0);
    for (int i = 0; i < ntile; i++) {
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      bar_wait(b_full + s, ph);
      bar_wait(acc_empty + s, ph ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; kk++) mma_i8(tmem + s * kN, da[kk], s ? db[1][kk] : db[0][kk], kk > 0);
        mma_commit(b_empty + s);   // the stage may be refilled once these MMAs have read it
        mma_commit(acc_full + s);  // ... and the accumulator is complete
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int row = q0 + quad * 32 + lane;
    uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
    int thr = -100000;  // accumulators above thr have a distance strictly below the current second best
    for (int i = 0; i < ntile; i++) {
      const int s = i & 1;
      bar_wait(acc_full + s, (i >> 1) & 1);
      tc_fence_after();
      const int col0 = (tb + i) * kN;
#pragma unroll 1
      for (int c0 = 0; c0 < kN / 32; c0 += kGroups) {
        int v[kGroups][32];
#pragma unroll
        for (int g = 0; g < kGroups; g++)
          tmem_ld32_issue(tmem + ((uint32_t)(quad * 32) << 16) + s * kN + (c0 + g) * 32, v[g]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < kGroups; g++) {
          const int c = c0 + g;
          bool hit = true;
          if (kFilter) {
            int mx = v[g][0];
#pragma unroll
            for (int j = 1; j < 31; j += 2) mx = max(mx, max(v[g][j], v[g][j + 1]));
            mx = max(mx, v[g][31]);
            hit = mx > thr;
          }
          if (hit) {
            const uint32_t base = (256u << 21) | (uint32_t)(col0 + c * 32);
#pragma unroll
            for (int j = 0; j < 32; j++) {
              if (col0 + c * 32 + j < nt) {  // rows past the end are TMA zero fill (accumulator 0 = distance 128)
                const uint32_t key = base + j - ((uint32_t)v[g][j] << 21);  // ((256 - acc) / 2) << 22 | col
                k2 = min(k2, max(k1, key));
                k1 = min(k1, key);
              }
            }
            thr = 256 - 2 * (int)(k2 >> 22);
          }
        }
      }
      tc_fence_before();
      bar_arrive(acc_empty + s);
    }
    if (row < nq) {
      const int i1 = k1 == 0xffffffffu ? -1 : (int)(k1 & 0x3fffff), i2 = k2 == 0xffffffffu ? -1 : (int)(k2 & 0x3fffff);
      partial[(size_t)blockIdx.y * nq + row] =
          make_int4(i1 < 0 ? 0x7fffffff : (int)(k1 >> 22), i1, i2 < 0 ? 0x7fffffff : (int)(k2 >> 22), i2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------
// Second form (the default): 256 queries per CTA, both query tiles in TENSOR MEMORY.
//
// What k_knn2_tc above is bound by is neither the tensor pipe nor the epilogue but the L2: every CTA streams the whole
// expanded train set through its shared memory, 64 KB per 1024 clk of MMAs = 64 B/clk/SM, 9.5 KB/clk for 148 SMs against
// the ~6.3 KB/clk the L2 slices deliver chip-wide (B300_MICROARCH.md "LTS throughput cap"; measured here: the query
// tile in tensor memory, four train tiles in flight and a second set of epilogue warps moved 1.70 ms to 1.70 / 1.70 /
// 1.61 ms). The bytes per MAC must come down, i.e. every train tile must meet more query rows:
//  * the query rows never change during a CTA's life, so they live in tensor memory: the epilogue threads build TWO
//    128-row tiles once, straight from the 32-byte descriptors (bit -> +-1 byte; no expanded query array in HBM, no TMA
//    of A), tcgen05.st them into 2 x 64 columns, and every MMA takes A from there ("tcgen05.mma [d], [a], b-desc");
//  * a train tile of N = 192 rows is multiplied with query tile 0 into accumulator 0 and with query tile 1 into
//    accumulator 1 (2 x 192 + 2 x 64 = 512 columns): the two accumulators ARE the double buffer — the epilogue reads
//    one while the tensor pipe fills the other — and a train byte now serves 256 query rows: 32 B/clk/SM;
//  * shared memory holds nothing but train tiles: four of them in flight (4 x 48 KB);
//  * two epilogue warps per TMEM lane quarter, each taking half of a tile's columns with its own best-two and filter
//    threshold per query row (merged at the very end): the accumulator is released as soon as its values are in
//    registers.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kM2 = 256;         // queries per CTA = two tcgen05 M = 128 tiles
constexpr int kN2 = 192, kB2Bytes = kN2 * kRowBytes, kChunks2 = kN2 / 32, kACol = 2 * kN2;
constexpr int kThreads2 = 320;   // TMA warp, MMA warp, 2 x 4 epilogue warps
constexpr int kChunksLo = kChunks2 / 2;  // chunks of a tile read by the first epilogue warp of a quarter; the rest by the second
constexpr int kStages2 = 4;      // train tiles in flight
constexpr int kSmem2 = kStages2 * kB2Bytes + 256 + 1024;
constexpr uint32_t kIdesc2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN2 >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
static_assert(kN2 % 32 == 0 && kACol + 128 <= 512, "TMEM budget");
static_assert(kSmem2 <= 227 * 1024, "shared memory");

__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(kIdesc2), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// 4 bits -> 4 bytes of +-1, the order of k_expand_pm1 (bit j -> byte j)
__device__ __forceinline__ uint32_t pm1_word(uint32_t nibble) {
  const uint32_t spread = (nibble & 1u) | ((nibble & 2u) << 7) | ((nibble & 4u) << 14) | ((nibble & 8u) << 21);  // 0/1 per byte
  return 0xffffffffu - spread * 0xfeu;  // byte = 0xff (bit clear: -1) or 0x01 (bit set: +1)
}

__global__ void __launch_bounds__(kThreads2, 1)
k_knn2_tc_ts(const uint8_t* __restrict__ q, const __grid_constant__ CUtensorMap map_t, int nq, int nt, int tiles_per_split,
             int4* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;  // [kStages2][2 halves][192 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages2 * kB2Bytes);
  uint64_t *a_full = bars, *b_full = bars + 1, *b_empty = b_full + kStages2, *acc_full = b_empty + kStages2,
           *acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int q0 = blockIdx.x * kM2;
  const int total_tiles = (nt + kN2 - 1) / kN2;
  const int tb = blockIdx.y * tiles_per_split;
  const int ntile = max(0, min(total_tiles, tb + tiles_per_split) - tb);

  if (threadIdx.x == 0) {
    bar_init(a_full, 128);
    for (int s = 0; s < kStages2; s++) {
      bar_init(b_full + s, 1);
      bar_init(b_empty + s, 1);
    }
    for (int s = 0; s < 2; s++) {
      bar_init(acc_full + s, 1);
      bar_init(acc_empty + s, 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // epilogue threads: best two keys of their two rows (query tile 0 / 1) over their half of the columns
  uint32_t k1[2] = {0xffffffffu, 0xffffffffu}, k2[2] = {0xffffffffu, 0xffffffffu};
  const int quad = warp & 3, half = (warp - 2) >> 2;  // warp w may touch TMEM lanes 32 (w % 4) .. +31
  const int row0 = q0 + quad * 32 + lane;             // the thread's row of query tile 0; tile 1: + 128

  if (warp == 0) {
    for (int i = 0; i < ntile; i++) {
      const int s = i % kStages2;
      bar_wait(b_empty + s, ((i / kStages2) & 1) ^ 1);
      if (elect_one()) {
        bar_expect(b_full + s, kB2Bytes);
        tma_rows(&map_t, sB + s * kB2Bytes, b_full + s, 0, (tb + i) * kN2);
        tma_rows(&map_t, sB + s * kB2Bytes + kN2 * kHalf, b_full + s, kHalf, (tb + i) * kN2);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    uint64_t db[8];  // stage 0; stage s starts s * kB2Bytes further on (the address field counts 16-byte units)
#pragma unroll
    for (int kk = 0; kk < 8; kk++) db[kk] = smem_desc(sB + (kk >> 2) * (kN2 * kHalf) + (kk & 3) * 32);
    bar_wait(a_full, 0);
    for (int i = 0; i < ntile; i++) {
      const int s = i % kStages2;
      bar_wait(b_full + s, (i / kStages2) & 1);
      const uint64_t soff = (uint64_t)(s * (kB2Bytes >> 4));
#pragma unroll
      for (int a = 0; a < 2; a++) {  // query tile a x train tile i -> accumulator a
        bar_wait(acc_empty + a, (i & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; kk++)  // K step kk = bytes 32 kk .. 32 kk + 31 of a row = A columns 8 kk .. 8 kk + 7
            mma_i8_ts(tmem + a * kN2, tmem + kACol + 64 * a + 8 * kk, db[kk] + soff, kk > 0);
          if (a == 1) mma_commit(b_empty + s);  // both query tiles have read the stage
          mma_commit(acc_full + a);
        }
        __syncwarp();
      }
    }
  } else {
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    if (half == 0) {
      // ---- the query tiles: lane = query row; descriptor byte b -> A columns 2 b (bits 0..3) and 2 b + 1 (bits 4..7),
      //      i.e. element 8 b + j = bit j of byte b, exactly the layout k_expand_pm1 gives the train rows ----
#pragma unroll 1
      for (int a = 0; a < 2; a++) {
        const int row = row0 + 128 * a;
        uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
        if (row < nq) {
          d0 = __ldg(reinterpret_cast<const uint4*>(q + (size_t)row * 32));
          d1 = __ldg(reinterpret_cast<const uint4*>(q + (size_t)row * 32) + 1);
        }
        const uint32_t w[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int h = 0; h < 2; h++) {
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 32; c++) {
            const int col = 32 * h + c;                     // column = 4 elements = one nibble of descriptor byte col / 2
            const uint32_t byte = (w[col >> 3] >> (8 * ((col >> 1) & 3))) & 0xffu;
            v[c] = pm1_word((col & 1) ? byte >> 4 : byte & 15u);
          }
          tmem_st32(lane_base + kACol + 64 * a + 32 * h, v);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      bar_arrive(a_full);
    }
    int thr[2] = {-100000, -100000};
    const int cfirst = half == 0 ? 0 : kChunksLo;  // this warp's chunks of every tile: cfirst .. cfirst + kChunksLo - 1
    for (int i = 0; i < ntile; i++) {
      const int col0 = (tb + i) * kN2 + cfirst * 32;
#pragma unroll
      for (int a = 0; a < 2; a++) {
        bar_wait(acc_full + a, i & 1);
        tc_fence_after();
        int v[kChunksLo][32];
#pragma unroll
        for (int g = 0; g < kChunksLo; g++) tmem_ld32_issue(lane_base + a * kN2 + (cfirst + g) * 32, v[g]);
        tmem_ld_wait();
        tc_fence_before();
        bar_arrive(acc_empty + a);  // the values are in registers: the accumulator may be overwritten
#pragma unroll
        for (int g = 0; g < kChunksLo; g++) {  // 32 accumulator columns = train rows cbase .. cbase + 31
          const int cbase = col0 + g * 32;
          int mx = v[g][0];
#pragma unroll
          for (int j = 1; j < 31; j += 2) mx = max(mx, max(v[g][j], v[g][j + 1]));
          mx = max(mx, v[g][31]);
          if (mx > thr[a]) {
            // exact insertion, a serial k1 / k2 chain over the 32 elements. (A tournament — the chunk's own best two by
            // independent pairwise merges, then one merge with the running pair — was measured SLOWER: 1.55 vs 1.52 ms
            // for 100k x 100k; the path is bound by instruction issue, not by the chain's latency.)
            const uint32_t base = (256u << 21) | (uint32_t)cbase;
#pragma unroll
            for (int j = 0; j < 32; j++) {
              if (cbase + j < nt) {  // rows past the end are TMA zero fill (accumulator 0 = distance 128)
                const uint32_t key = base + j - ((uint32_t)v[g][j] << 21);  // ((256 - acc) / 2) << 22 | col
                k2[a] = min(k2[a], max(k1[a], key));
                k1[a] = min(k1[a], key);
              }
            }
            thr[a] = 256 - 2 * (int)(k2[a] >> 22);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  // ---- merge the two column halves of every row (shared memory is free: every train tile has been consumed) ----
  uint4* xch = reinterpret_cast<uint4*>(smem);
  if (warp >= 2 && half == 1) xch[quad * 32 + lane] = make_uint4(k1[0], k2[0], k1[1], k2[1]);
  __syncthreads();
  if (warp >= 2 && half == 0) {
    const uint4 o = xch[quad * 32 + lane];
    const uint32_t o1[2] = {o.x, o.z}, o2[2] = {o.y, o.w};
#pragma unroll
    for (int a = 0; a < 2; a++) {
      const int row = row0 + 128 * a;
      if (row >= nq) continue;
      const uint32_t m1 = min(k1[a], o1[a]), m2 = min(max(k1[a], o1[a]), min(k2[a], o2[a]));
      const int i1 = m1 == 0xffffffffu ? -1 : (int)(m1 & 0x3fffff), i2 = m2 == 0xffffffffu ? -1 : (int)(m2 & 0x3fffff);
      partial[(size_t)blockIdx.y * nq + row] =
          make_int4(i1 < 0 ? 0x7fffffff : (int)(m1 >> 22), i1, i2 < 0 ? 0x7fffffff : (int)(m2 >> 22), i2);
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encoder() {
  static EncodeTiledFn enc = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return enc;
}

// rows x 256 signed bytes, box = 128 bytes (one swizzle span) x box_rows; rows past the end read as zero
bool make_rows_map(CUtensorMap* m, const void* base, int rows, int box_rows) {
  EncodeTiledFn enc = encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)kRowBytes, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kRowBytes};
  const cuuint32_t box[2] = {(cuuint32_t)kHalf, (cuuint32_t)box_rows}, es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// Large enough for the GEMM form to pay for its expansion pass and 128-row query tiles; keys need nt < 2^22 and the
// expanded train set (256 B per row) is kept to 256 MB of scratch.
bool knn2_tc_eligible(int nq, int nt) {
  return nq >= 1024 && nq <= (1 << 24) && nt >= 256 && nt <= (1 << 20) && (long long)nq * nt >= (1ll << 25) && encoder() != nullptr;
}

size_t knn2_tc_expanded_bytes(int rows) { return (size_t)rows * kRowBytes; }
static bool knn2_ts();
bool knn2_tc_expands_queries() { return !knn2_ts(); }

// Splits of the train set. Measured on B200 (tools/ubench/knn2_tc.cu, round 2): a split costs more than it evens out —
// every CTA reloads its query tile, and its chunk filter starts cold, so the first tiles of every split take the exact
// insertion path (100k x 100k: 1.77 / 1.77 / 1.84 / 2.00 / 2.26 ms for 1 / 2 / 3 / 4 / 7 splits; 50k x 70k: 0.67 / 0.73 /
// 0.72 / 0.78; 30k x 30k: 0.23 / 0.26 / 0.28 / 0.32). Splitting only pays while the query blocks alone leave SMs idle
// (10k x 10k, 79 blocks: 0.074 / 0.070 / 0.072 / 0.078 ms for 2 / 4 / 8 / 14). So: one split once there is a CTA per SM,
// otherwise enough to put ~2 CTAs on every SM; every split keeps at least 4 tiles.
// ORBM_KNN2_TS=0 keeps the query tile in shared memory (k_knn2_tc); default: tensor memory (k_knn2_tc_ts)
static bool knn2_ts() {
  static const bool on = [] {
    const char* e = getenv("ORBM_KNN2_TS");
    return !(e && e[0] == '0');
  }();
  return on;
}

int knn2_tc_splits(int nq, int nt, int* tiles_per_split) {
  if (knn2_ts()) {
    // k_knn2_tc_ts: rounds of 148 CTAs x (tiles per CTA + a fixed cost per CTA: prologue, first TMA round trip and the
    // cold chunk filter of its first tiles; measured ~42 tile times: 100k x 100k takes 1.52 ms in one split, 1.56 ms in
    // three although three fill the last round of CTAs) — the split count with the least total wins
    const int qblocks = (nq + kM2 - 1) / kM2, total_tiles = (nt + kN2 - 1) / kN2;
    static const int c0 = getenv("ORBM_KNN2_C0") ? atoi(getenv("ORBM_KNN2_C0")) : 45;  // measured: ~42 (tuning aid)
    int best_s = 1;
    long long best = -1;
    for (int s = 1; s <= 16 && (s == 1 || total_tiles / s >= 4); s++) {
      const int tps = (total_tiles + s - 1) / s, real = (total_tiles + tps - 1) / tps;
      const long long rounds = ((long long)qblocks * real + 147) / 148;
      const long long cost = rounds * (tps + c0);
      if (best < 0 || cost < best) {
        best = cost;
        best_s = s;
      }
    }
    const int tps = (total_tiles + best_s - 1) / best_s;
    *tiles_per_split = tps;
    return (total_tiles + tps - 1) / tps;
  }
  const int qblocks = (nq + kM - 1) / kM, total_tiles = (nt + kN - 1) / kN;
  int s = 1;
  if (qblocks < 148) {
    s = (2 * 148 + qblocks - 1) / qblocks;
    if (s > 16) s = 16;
    if (s > total_tiles / 4) s = total_tiles / 4;
    if (s < 1) s = 1;
  }
  const int tps = (total_tiles + s - 1) / s;
  *tiles_per_split = tps;
  return (total_tiles + tps - 1) / tps;  // no empty split
}

cudaError_t launch_knn2_tc(const uint8_t* q, int nq, const uint8_t* t, int nt, int8_t* expanded_q, int8_t* expanded_t,
                           int4* partial, int splits, int tiles_per_split, cudaStream_t st) {
  CUtensorMap mq, mt;
  if (knn2_ts()) {  // the query tile is built in tensor memory from the raw descriptors: no expanded query array
    if (!make_rows_map(&mt, expanded_t, nt, kN2)) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_knn2_tc_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2);
    if (e != cudaSuccess) return e;
    k_expand_pm1<<<(nt * 32 + 255) / 256, 256, 0, st>>>(t, nt, expanded_t);
    k_knn2_tc_ts<<<dim3((nq + kM2 - 1) / kM2, splits), kThreads2, kSmem2, st>>>(q, mt, nq, nt, tiles_per_split, partial);
    return cudaGetLastError();
  }
  if (!make_rows_map(&mq, expanded_q, nq, kM) || !make_rows_map(&mt, expanded_t, nt, kN)) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_knn2_tc<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  if (e != cudaSuccess) return e;
  k_expand_pm1<<<(nq * 32 + 255) / 256, 256, 0, st>>>(q, nq, expanded_q);
  k_expand_pm1<<<(nt * 32 + 255) / 256, 256, 0, st>>>(t, nt, expanded_t);
  k_knn2_tc<true, 4><<<dim3((nq + kM - 1) / kM, splits), kThreads, kSmem, st>>>(mq, mt, nq, nt, tiles_per_split, partial);
  return cudaGetLastError();
}

}  // namespace orbx
