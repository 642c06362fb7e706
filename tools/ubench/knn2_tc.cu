// knn2_tc — stand-alone form of the tensor-core variant of orbm_knn2: checks itself against a
// brute-force POPC kernel and prints the rate. Not linked into the library.
//
//   STATUS: ran on a B200 at the end of round 1 (profiles/r01_knn2_tc_prototype.log): 0 of 100 000 rows differ from the
//   brute-force kernel, 100k x 100k in 2.20 ms (3.66 ms with the chunk filter off). The kernel now lives in the library
//   as orb_slam3_fast_b200/csrc/k_knn2_tc.cu; this file stays as the stand-alone bench / bring-up harness:
//       nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o knn2_tc knn2_tc.cu -lcuda
//       timeout 120 ./knn2_tc 1000 1000 && timeout 120 ./knn2_tc 100000 100000     # [nq] [nt] [splits] [filter 0|1] [loads in flight 1|2|4]
//
// Idea. A 256-bit descriptor is expanded once to 256 signed bytes of +-1 (k_expand_pm1, 32 B -> 256 B per row); then
//   dot(a', b') = 256 - 2 * hamming(a, b)      exactly, in int32,
// which is a plain s8 x s8 -> s32 GEMM: tcgen05.mma kind::i8, M = 128 queries x N = 256 train rows x K = 256 per tile,
// operands K-major in 128B-swizzled shared memory written by TMA, the accumulator double-buffered in TMEM (2 x 256
// columns), so the epilogue of tile i overlaps the MMAs of tile i + 1. The epilogue owns one query row per thread
// (TMEM lane), reads 32 accumulator columns per tcgen05.ld and keeps the best two as packed (distance << 22 | train row)
// keys — the same keys, and therefore the same "lower trainIdx wins ties" order, as k_knn2 in csrc/k_match.cu. Because
// train tiles are visited in ascending row order a candidate can only enter the best two if its distance is STRICTLY
// below the current second best, so a 32-column chunk is skipped after one 3-input max reduction (0.5 op per value)
// unless its largest accumulator beats that threshold; the exact insertion (5 ops per value) runs on ~2 ln(nt) chunks
// per row. Budget per 128 x 256 tile and SM: MMA 1024 clk at the int8 peak, epilogue ~300 clk on the fast path —
// 100k x 100k = 305 k tiles -> ~1.1 ms on 148 SMs against 15.0 ms for the POPC kernel.
//
// One CTA per SM (160 KB of shared memory, all 512 TMEM columns): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = epilogue (warp w may touch TMEM lanes 32 (w % 4) .. +31).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

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
// Bounded spin: a wrong byte count or parity in a first run must end in a trap, not in a hung GPU box.
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  for (uint32_t spins = 0;; spins++) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(sptr(b)), "r"(parity) : "memory");
    if (ok) return;
    if (spins > (1u << 24)) {
      printf("knn2_tc: barrier %d of block (%d,%d) thread %d never completed (parity %u)\n",
             (int)(sptr(b) & 255) / 8, blockIdx.x, blockIdx.y, threadIdx.x, parity);
      __trap();
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

template <bool kFilter, int kGroups>  // kGroups = 32-column TMEM loads in flight per tcgen05.wait::ld
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      bar_expect(a_full, kABytes);
      tma_rows(&map_q, sA, a_full, 0, q0);
      tma_rows(&map_q, sA + kM * kHalf, a_full, kHalf, q0);
      for (int i = 0; i < ntile; i++) {
        const int s = i & 1;
        bar_wait(b_empty + s, ((i >> 1) & 1) ^ 1);
        bar_expect(b_full + s, kBBytes);
        tma_rows(&map_t, sB + s * kBBytes, b_full + s, 0, (tb + i) * kN);
        tma_rows(&map_t, sB + s * kBBytes + kN * kHalf, b_full + s, kHalf, (tb + i) * kN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      bar_wait(a_full, 0);
      for (int i = 0; i < ntile; i++) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        bar_wait(b_full + s, ph);
        bar_wait(acc_empty + s, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {  // 8 x (K = 32 bytes); 4 steps inside each 128-byte swizzle span
          const uint64_t da = smem_desc(sA + (kk >> 2) * (kM * kHalf) + (kk & 3) * 32);
          const uint64_t db = smem_desc(sB + s * kBBytes + (kk >> 2) * (kN * kHalf) + (kk & 3) * 32);
          mma_i8(tmem + s * kN, da, db, kk > 0);
        }
        mma_commit(b_empty + s);   // the stage may be refilled once these MMAs have read it
        mma_commit(acc_full + s);  // ... and the accumulator is complete
      }
    }
    __syncwarp();
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

// brute force checker: thread = query, every train row, POPC
__global__ void k_knn2_check(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt,
                             int4* __restrict__ out) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  uint32_t a[8];
  for (int k = 0; k < 8; k++) a[k] = reinterpret_cast<const uint32_t*>(q)[(size_t)qi * 8 + k];
  uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
  for (int r = 0; r < nt; r++) {
    int d = 0;
    for (int k = 0; k < 8; k++) d += __popc(a[k] ^ __ldg(reinterpret_cast<const uint32_t*>(t) + (size_t)r * 8 + k));
    const uint32_t key = ((uint32_t)d << 22) | (uint32_t)r;
    k2 = min(k2, max(k1, key));
    k1 = min(k1, key);
  }
  const int i1 = k1 == 0xffffffffu ? -1 : (int)(k1 & 0x3fffff), i2 = k2 == 0xffffffffu ? -1 : (int)(k2 & 0x3fffff);
  out[qi] = make_int4(i1 < 0 ? 0x7fffffff : (int)(k1 >> 22), i1, i2 < 0 ? 0x7fffffff : (int)(k2 >> 22), i2);
}

__global__ void k_merge(const int4* __restrict__ partial, int nq, int splits, int4* __restrict__ out) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
  for (int s = 0; s < splits; s++) {
    const int4 p = partial[(size_t)s * nq + qi];
    const uint32_t c[2] = {p.y < 0 ? 0xffffffffu : ((uint32_t)p.x << 22) | (uint32_t)p.y,
                           p.w < 0 ? 0xffffffffu : ((uint32_t)p.z << 22) | (uint32_t)p.w};
    for (int k = 0; k < 2; k++) {
      k2 = min(k2, max(k1, c[k]));
      k1 = min(k1, c[k]);
    }
  }
  const int i1 = k1 == 0xffffffffu ? -1 : (int)(k1 & 0x3fffff), i2 = k2 == 0xffffffffu ? -1 : (int)(k2 & 0x3fffff);
  out[qi] = make_int4(i1 < 0 ? 0x7fffffff : (int)(k1 >> 22), i1, i2 < 0 ? 0x7fffffff : (int)(k2 >> 22), i2);
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_map(Enc enc, CUtensorMap* m, void* base, int rows, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)kRowBytes, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kRowBytes};
  const cuuint32_t box[2] = {(cuuint32_t)kHalf, (cuuint32_t)box_rows}, es[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));        \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)
}  // namespace

int main(int argc, char** argv) {
  const int nq = argc > 1 ? atoi(argv[1]) : 10000, nt = argc > 2 ? atoi(argv[2]) : 10000;
  int splits = argc > 3 ? atoi(argv[3]) : 0;
  const bool filter = argc > 4 ? atoi(argv[4]) != 0 : true;
  const int groups = argc > 5 ? atoi(argv[5]) : 1;  // 1, 2 or 4 TMEM loads in flight
  if (nt >= (1 << 22) || nq < 1 || nt < 1) return printf("sizes out of range\n"), 1;
  const int qblocks = (nq + kM - 1) / kM, total_tiles = (nt + kN - 1) / kN;
  if (splits <= 0) splits = (2 * 148 + qblocks - 1) / qblocks;
  if (splits > total_tiles) splits = total_tiles;
  const int tiles_per_split = (total_tiles + splits - 1) / splits;
  splits = (total_tiles + tiles_per_split - 1) / tiles_per_split;  // no empty split

  // descriptors: random, with a quarter of the train rows drawn from 64 prototypes so that ties are common
  std::vector<uint8_t> hq((size_t)nq * 32), ht((size_t)nt * 32);
  uint64_t s = 0x9e3779b97f4a7c15ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (auto& b : hq) b = (uint8_t)(rnd() >> 24);
  for (auto& b : ht) b = (uint8_t)(rnd() >> 24);
  for (int r = 0; r < nt; r += 4) memcpy(&ht[(size_t)r * 32], &ht[(size_t)((r / 4) % 64) * 4 * 32], 32);
  for (int r = 0; r < nq; r += 7) memcpy(&hq[(size_t)r * 32], &ht[(size_t)(r % nt) * 32], 32);  // exact hits, d = 0

  uint8_t *dq, *dt;
  int8_t *eq, *et;
  int4 *partial, *out, *ref;
  CK(cudaMalloc(&dq, hq.size()));
  CK(cudaMalloc(&dt, ht.size()));
  CK(cudaMalloc(&eq, (size_t)nq * kRowBytes));
  CK(cudaMalloc(&et, (size_t)nt * kRowBytes));
  CK(cudaMalloc(&partial, (size_t)splits * nq * sizeof(int4)));
  CK(cudaMalloc(&out, (size_t)nq * sizeof(int4)));
  CK(cudaMalloc(&ref, (size_t)nq * sizeof(int4)));
  CK(cudaMemcpy(dq, hq.data(), hq.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dt, ht.data(), ht.size(), cudaMemcpyHostToDevice));

  void* p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
  CUtensorMap mq, mt;
  if (!p || !make_map((Enc)p, &mq, eq, nq, kM) || !make_map((Enc)p, &mt, et, nt, kN)) return 1;

  auto kern = !filter ? k_knn2_tc<false, 1> : groups == 4 ? k_knn2_tc<true, 4> : groups == 2 ? k_knn2_tc<true, 2> : k_knn2_tc<true, 1>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  auto run = [&]() {
    k_expand_pm1<<<(nq * 32 + 255) / 256, 256>>>(dq, nq, eq);
    k_expand_pm1<<<(nt * 32 + 255) / 256, 256>>>(dt, nt, et);
    kern<<<dim3(qblocks, splits), kThreads, kSmem>>>(mq, mt, nq, nt, tiles_per_split, partial);
    k_merge<<<(nq + 255) / 256, 256>>>(partial, nq, splits, out);
  };
  run();
  CK(cudaDeviceSynchronize());
  k_knn2_check<<<(nq + 127) / 128, 128>>>(dq, nq, dt, nt, ref);
  CK(cudaDeviceSynchronize());
  std::vector<int4> ho(nq), hr(nq);
  CK(cudaMemcpy(ho.data(), out, (size_t)nq * sizeof(int4), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hr.data(), ref, (size_t)nq * sizeof(int4), cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < nq; i++)
    if (memcmp(&ho[i], &hr[i], sizeof(int4)) != 0 && bad++ < 8)
      printf("row %d: tc (%d,%d,%d,%d) != check (%d,%d,%d,%d)\n", i, ho[i].x, ho[i].y, ho[i].z, ho[i].w, hr[i].x,
             hr[i].y, hr[i].z, hr[i].w);
  printf("parity: %ld of %d rows differ\n", bad, nq);

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) run();
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) run();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  printf("knn2_tc %d x %d splits %d filter %d groups %d: %.3f ms (expand + mma + merge) = %.1f Gpair/s\n", nq, nt,
         splits, (int)filter, groups, ms, (double)nq * nt / ms * 1e-6);
  return bad ? 2 : 0;
}
