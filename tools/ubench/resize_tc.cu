// resize_tc — stand-alone bring-up / bench harness of a tensor-core form of cv::resize(u8, INTER_LINEAR) as
// ORBextractor::ComputePyramid uses it (src/ORBextractor.cc:1108-1145; OpenCV's fixed point: 11-bit coefficients,
// horizontal c0*s0 + c1*s1, then ((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2 vertically). Checks itself
// against a plain per-pixel kernel and prints the rate.
//
//   timeout 120 ./resize_tc [frames=256] [w=640] [h=480] [levels=8] [frames per CTA=32]
//
// The HORIZONTAL pass is a product with a 2-banded coefficient matrix, i.e. a u8 x u8 -> s32 GEMM once the 11-bit
// coefficients are split into a low byte and a high part (<= 8): H = 256 * (A_hi x S) + A_lo x S, two accumulators of
// 6 MMAs each (K = 192 source columns cover 128 destination columns at scales up to ~1.37). The vertical pass floors
// each of its two terms separately, so it is not linear: it stays on the CUDA cores, fed from a u16 tile (H >> 4) in
// shared memory, 4 destination pixels per thread. warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = D -> H16,
// warps 6..13 = vertical pass + output tile + TMA store.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../orb_slam3_fast_b200/csrc/orbx_plan.h"

namespace {
struct Tab {  // = orbx::ResizeTab
  int16_t ofs, c0, c1, pad;
};
constexpr int kSlab = 4096;
constexpr int kK = 6;                       // K steps of 32 source columns
constexpr int kThreads = 14 * 32;
constexpr int kStageBytes = 16384 + 2 * kSlab;
constexpr int oAlo = 0;
constexpr int oAhi = oAlo + kK * kSlab;
constexpr int oS = oAhi + kK * kSlab;       // 2 stages
constexpr int oH = oS + 2 * kStageBytes;    // 2 x [128 rows][128 x] u16
constexpr int oOut = oH + 2 * 128 * 256;    // 2 x [128 rows][128 B], 128B-swizzled
constexpr int oTabs = oOut + 2 * 128 * 128; // per tile row: sy0, sy1 (u8), b0 << 16, b1 << 16 (u32)
constexpr int oBar = oTabs + 128 * 12;
constexpr int kSmem = oBar + 256 + 1024;
static_assert(kSmem <= 232448, "shared memory");
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // u8 x u8 -> s32

struct Params {
  CUtensorMap src;   // box 128 B x 128 rows, SWIZZLE_128B
  CUtensorMap src2;  // box 32 B x 128 rows, SWIZZLE_32B
  CUtensorMap dst;   // box 128 B x yo rows, SWIZZLE_128B
  const Tab* xtab;
  const Tab* ytab;
  int sw, sh, dw, dh, yo, frames, fpc;
};

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
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  for (uint32_t spins = 0;; spins++) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(sptr(b)), "r"(parity) : "memory");
    if (ok) return;
    if (spins > (1u << 22)) {
      printf("resize_tc: barrier at %u of block (%d,%d) thread %d never completed (parity %u)\n", sptr(b), blockIdx.x,
             blockIdx.y, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load3(const CUtensorMap* map, void* dst, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(sptr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(sptr(bar)) : "memory");
}
__device__ __forceinline__ void tma_store3(const CUtensorMap* map, const void* src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(sptr(src)) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ uint64_t smem_desc32(const void* p) {
  const uint64_t a = (sptr(p) & 0x3ffff) >> 4;
  return a | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}
__device__ __forceinline__ uint64_t smem_desc128(const void* p) {
  const uint64_t a = (sptr(p) & 0x3ffff) >> 4;
  return a | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_u8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sptr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ int sw32(int row, int k) {
  return (k >> 5) * kSlab + (row >> 3) * 256 + (row & 7) * 32 + ((((k >> 4) & 1) ^ ((row >> 2) & 1)) << 4) + (k & 15);
}

__global__ void __launch_bounds__(kThreads, 1) k_resize_tc(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sAlo = smem + oAlo;
  uint8_t* sAhi = smem + oAhi;
  uint8_t* sS = smem + oS;
  uint8_t* sH = smem + oH;
  uint8_t* sOut = smem + oOut;
  uint8_t* sTabs = smem + oTabs;  // [128] sy0 | [128] sy1 | [128] u32 b0 << 16 | [128] u32 b1 << 16
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);
  uint64_t *s_full = bars, *s_empty = bars + 2, *d_full = bars + 4, *d_empty = bars + 6, *h_full = bars + 8,
           *h_empty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int ntx = (P.dw + 127) / 128;
  const int ty = blockIdx.x / ntx, tx = blockIdx.x - ty * ntx;
  const int x0 = tx * 128, y0 = ty * P.yo;
  const int f0 = blockIdx.y * P.fpc, n = min(P.frames, f0 + P.fpc) - f0;
  const int sh1 = P.sh - 1;
  auto clip = [&](int v) { return v < 0 ? 0 : (v > sh1 ? sh1 : v); };
  const int xa = P.xtab[x0].ofs & ~15;
  const int ys0 = clip(P.ytab[y0].ofs);
  const int rows = min(P.yo, P.dh - y0);  // destination rows of this tile

  if (tid == 0) {
    for (int i = 0; i < 2; i++) {
      bar_init(s_full + i, 1);
      bar_init(s_empty + i, 1);
      bar_init(d_full + i, 1);
      bar_init(d_empty + i, 128);
      bar_init(h_full + i, 128);
      bar_init(h_empty + i, 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * kK * kSlab / 16; i += kThreads) reinterpret_cast<uint4*>(sAlo)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid < 128) {
    const int x = x0 + tid;
    if (x < P.dw) {
      const Tab e = P.xtab[x];
      const int k = e.ofs - xa;
      sAlo[sw32(tid, k)] = (uint8_t)(e.c0 & 255);
      sAhi[sw32(tid, k)] = (uint8_t)(e.c0 >> 8);
      if (e.c1 != 0) {  // weight 0 on the clamped last column: its tap may lie outside the tile
        sAlo[sw32(tid, k + 1)] = (uint8_t)(e.c1 & 255);
        sAhi[sw32(tid, k + 1)] = (uint8_t)(e.c1 >> 8);
      }
    }
  } else if (tid < 256) {
    const int r = tid - 128;
    if (r < rows) {
      const Tab e = P.ytab[y0 + r];
      sTabs[r] = (uint8_t)(clip(e.ofs) - ys0);
      sTabs[128 + r] = (uint8_t)(clip(e.ofs + 1) - ys0);
      reinterpret_cast<uint32_t*>(sTabs + 256)[r] = (uint32_t)e.c0 << 16;
      reinterpret_cast<uint32_t*>(sTabs + 768)[r] = (uint32_t)e.c1 << 16;
    }
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    for (int i = 0; i < n; i++) {
      const int st = i & 1;
      bar_wait(s_empty + st, ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* stage = sS + st * kStageBytes;
        bar_expect(s_full + st, kStageBytes);
        tma_load3(&P.src, stage, s_full + st, xa, ys0, f0 + i);
        tma_load3(&P.src2, stage + 16384, s_full + st, xa + 128, ys0, f0 + i);
        tma_load3(&P.src2, stage + 16384 + kSlab, s_full + st, xa + 160, ys0, f0 + i);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    uint64_t dAlo[kK], dAhi[kK], dS[2][kK];
#pragma unroll
    for (int s = 0; s < kK; s++) {
      dAlo[s] = smem_desc32(sAlo + s * kSlab);
      dAhi[s] = smem_desc32(sAhi + s * kSlab);
#pragma unroll
      for (int st = 0; st < 2; st++) {
        const uint8_t* stage = sS + st * kStageBytes;
        dS[st][s] = s < 4 ? smem_desc128(stage + 32 * s) : smem_desc32(stage + 16384 + (s - 4) * kSlab);
      }
    }
    for (int i = 0; i < n; i++) {
      const int st = i & 1;
      bar_wait(s_full + st, (i >> 1) & 1);
      bar_wait(d_empty + st, ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < kK; s++) mma_u8(tmem + st * 256, dAlo[s], st ? dS[1][s] : dS[0][s], s > 0);
#pragma unroll
        for (int s = 0; s < kK; s++) mma_u8(tmem + st * 256 + 128, dAhi[s], st ? dS[1][s] : dS[0][s], s > 0);
        mma_commit(s_empty + st);
        mma_commit(d_full + st);
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ---- D (lanes = destination column x, columns = source row r) -> H16[r][x] = H >> 4 ----
    const int x = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < n; i++) {
      const int b = i & 1;
      uint16_t* h16 = reinterpret_cast<uint16_t*>(sH + b * 128 * 256) + x;
      bar_wait(d_full + b, (i >> 1) & 1);
      tc_fence_after();
      bar_wait(h_empty + b, ((i >> 1) & 1) ^ 1);
#pragma unroll 1
      for (int c = 0; c < 4; c++) {
        uint32_t lo[32], hi[32];
        tmem_ld32(lane_base + b * 256 + c * 32, lo);
        tmem_ld32(lane_base + b * 256 + 128 + c * 32, hi);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) h16[(c * 32 + j) * 128] = (uint16_t)((hi[j] << 4) + (lo[j] >> 4));  // (256 hi + lo) >> 4
      }
      tc_fence_before();
      bar_arrive(d_empty + b);
      bar_arrive(h_full + b);
    }
  } else {
    // ---- vertical pass: 4 destination pixels per thread, rows y = rw, rw + 8, ... ----
    const int tv = tid - 6 * 32, xg = tv & 31, rw = tv >> 5;
    const bool leader = tv == 0;
    const uint32_t* b0s = reinterpret_cast<const uint32_t*>(sTabs + 256);
    const uint32_t* b1s = reinterpret_cast<const uint32_t*>(sTabs + 768);
    for (int i = 0; i < n; i++) {
      const int b = i & 1;
      const uint8_t* hb = sH + b * 128 * 256 + xg * 8;
      uint8_t* ob = sOut + b * 128 * 128;
      if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bar_wait(h_full + b, (i >> 1) & 1);
      for (int y = rw; y < rows; y += 8) {
        const uint2 p0 = *reinterpret_cast<const uint2*>(hb + (int)sTabs[y] * 256);
        const uint2 p1 = *reinterpret_cast<const uint2*>(hb + (int)sTabs[128 + y] * 256);
        const uint32_t c0 = b0s[y], c1 = b1s[y];
        const uint32_t v0 = (__umulhi(c0, p0.x & 0xffffu) + __umulhi(c1, p1.x & 0xffffu) + 2u) >> 2;
        const uint32_t v1 = (__umulhi(c0, p0.x >> 16) + __umulhi(c1, p1.x >> 16) + 2u) >> 2;
        const uint32_t v2 = (__umulhi(c0, p0.y & 0xffffu) + __umulhi(c1, p1.y & 0xffffu) + 2u) >> 2;
        const uint32_t v3 = (__umulhi(c0, p0.y >> 16) + __umulhi(c1, p1.y >> 16) + 2u) >> 2;
        const uint32_t word = __byte_perm(__byte_perm(v0, v1, 0x0040), __byte_perm(v2, v3, 0x0040), 0x5410);
        *reinterpret_cast<uint32_t*>(ob + y * 128 + (((xg >> 2) ^ (y & 7)) << 4) + (xg & 3) * 4) = word;
      }
      bar_arrive(h_empty + b);
      proxy_fence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (leader) tma_store3(&P.dst, ob, x0, y0, f0 + i);
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// plain reference: one thread per destination pixel (the arithmetic of k_resize's byte path)
__global__ void k_resize_ref(const uint8_t* src, int sw, int sh, int spitch, int64_t sfstride, uint8_t* dst, int dw, int dh,
                             int dpitch, int64_t dfstride, const Tab* xtab, const Tab* ytab) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (x >= dw) return;
  const Tab tx = xtab[x], ty = ytab[y];
  const int sh1 = sh - 1, sw1 = sw - 1;
  const int a0 = ty.ofs < 0 ? 0 : (ty.ofs > sh1 ? sh1 : ty.ofs);
  const int a1 = ty.ofs + 1 < 0 ? 0 : (ty.ofs + 1 > sh1 ? sh1 : ty.ofs + 1);
  const int s1 = tx.ofs + 1 < sw1 ? tx.ofs + 1 : sw1;
  const uint8_t* s = src + f * sfstride;
  const int h0 = ((int)s[(int64_t)a0 * spitch + tx.ofs] * tx.c0 + (int)s[(int64_t)a0 * spitch + s1] * tx.c1) >> 4;
  const int h1 = ((int)s[(int64_t)a1 * spitch + tx.ofs] * tx.c0 + (int)s[(int64_t)a1 * spitch + s1] * tx.c1) >> 4;
  dst[f * dfstride + (int64_t)y * dpitch + x] = (uint8_t)((((ty.c0 * h0) >> 16) + ((ty.c1 * h1) >> 16) + 2) >> 2);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)
}  // namespace

int main(int argc, char** argv) {
  const int frames = argc > 1 ? atoi(argv[1]) : 256;
  const int W = argc > 2 ? atoi(argv[2]) : 640, H = argc > 3 ? atoi(argv[3]) : 480;
  const int nl = argc > 4 ? atoi(argv[4]) : 8;
  const int fpc = argc > 5 ? atoi(argv[5]) : 32;
  const float scale = argc > 6 ? (float)atof(argv[6]) : 1.2f;
  int w[16], h[16], pitch[16];
  int64_t off[16], total = 0;
  float sf = 1.f;
  for (int l = 0; l < nl; l++) {
    w[l] = orbx::cv_round((float)W * (1.0f / sf));
    h[l] = orbx::cv_round((float)H * (1.0f / sf));
    sf = (float)(sf * (double)scale);
    pitch[l] = (w[l] + 63) / 64 * 64;
    off[l] = total;
    total += (int64_t)pitch[l] * h[l];
  }
  const int64_t fstride = (total + 255) / 256 * 256;
  uint8_t *pyr, *ref;
  CK(cudaMalloc(&pyr, fstride * frames));
  CK(cudaMalloc(&ref, fstride * frames));
  {
    std::vector<uint8_t> hs((size_t)fstride * frames);
    uint32_t s = 777;
    for (size_t i = 0; i < hs.size(); i++) {
      s = s * 1664525u + 1013904223u;
      hs[i] = (i / 911) % 9 == 0 ? 255 : ((i / 1201) % 13 == 0 ? 0 : (uint8_t)(s >> 24));
    }
    CK(cudaMemcpy(pyr, hs.data(), hs.size(), cudaMemcpyHostToDevice));  // every level starts as noise: each level's
    CK(cudaMemcpy(ref, hs.data(), hs.size(), cudaMemcpyHostToDevice));  // source is what the previous step left there
  }
  // tables
  std::vector<Tab> tabs;
  int xoff[16], yoff[16];
  for (int l = 1; l < nl; l++) {
    for (int axis = 0; axis < 2; axis++) {
      const int ss = axis ? h[l - 1] : w[l - 1], ds = axis ? h[l] : w[l];
      std::vector<int16_t> o(ds), c0(ds), c1(ds);
      orbx::axis_table(ss, ds, axis == 0, o.data(), c0.data(), c1.data());
      (axis ? yoff : xoff)[l] = (int)tabs.size();
      for (int d = 0; d < ds; d++) tabs.push_back(Tab{o[d], c0[d], c1[d], 0});
    }
  }
  Tab* d_tabs;
  CK(cudaMalloc(&d_tabs, tabs.size() * sizeof(Tab)));
  CK(cudaMemcpy(d_tabs, tabs.data(), tabs.size() * sizeof(Tab), cudaMemcpyHostToDevice));

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  Params P[16];
  dim3 grids[16];
  bool ok_level[16];
  for (int l = 1; l < nl; l++) {
    Params& p = P[l];
    memset(&p, 0, sizeof(p));
    p.xtab = d_tabs + xoff[l];
    p.ytab = d_tabs + yoff[l];
    p.sw = w[l - 1]; p.sh = h[l - 1]; p.dw = w[l]; p.dh = h[l];
    p.frames = frames; p.fpc = fpc;
    // the largest tile height whose source rows fit the 128-row box, and the column span check
    const Tab* xt = tabs.data() + xoff[l];
    const Tab* yt = tabs.data() + yoff[l];
    auto clip = [&](int v) { return v < 0 ? 0 : (v > p.sh - 1 ? p.sh - 1 : v); };
    int yo = 120;
    for (; yo >= 16; yo -= 8) {
      bool fits = true;
      for (int y0 = 0; y0 < p.dh && fits; y0 += yo) {
        const int y1 = std::min(y0 + yo, p.dh) - 1;
        fits = clip(yt[y1].ofs + 1) - clip(yt[y0].ofs) + 1 <= 128;
      }
      if (fits) break;
    }
    bool xfits = true;
    for (int x0 = 0; x0 < p.dw; x0 += 128) {
      const int x1 = std::min(x0 + 128, p.dw) - 1;
      const int xa = xt[x0].ofs & ~15;
      for (int x = x0; x <= x1; x++) {
        const int k = xt[x].ofs - xa + (xt[x].c1 != 0 ? 1 : 0);
        if (k >= 192 || xt[x].ofs - xa < 0) xfits = false;
      }
    }
    ok_level[l] = xfits && yo >= 16;
    p.yo = yo;
    const cuuint64_t sdims[3] = {(cuuint64_t)p.sw, (cuuint64_t)p.sh, (cuuint64_t)frames};
    const cuuint64_t sstr[2] = {(cuuint64_t)pitch[l - 1], (cuuint64_t)fstride};
    const cuuint64_t ddims[3] = {(cuuint64_t)p.dw, (cuuint64_t)p.dh, (cuuint64_t)frames};
    const cuuint64_t dstr[2] = {(cuuint64_t)pitch[l], (cuuint64_t)fstride};
    const cuuint32_t es[3] = {1, 1, 1};
    const cuuint32_t b1[3] = {128, 128, 1}, b2[3] = {32, 128, 1}, b3[3] = {128, (cuuint32_t)yo, 1};
    CUresult r1 = enc(&p.src, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr + off[l - 1], sdims, sstr, b1, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&p.src2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr + off[l - 1], sdims, sstr, b2, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r3 = enc(&p.dst, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr + off[l], ddims, dstr, b3, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 || r2 || r3) {
      printf("tensor map encode failed at level %d: %d %d %d\n", l, (int)r1, (int)r2, (int)r3);
      return 2;
    }
    grids[l] = dim3(((p.dw + 127) / 128) * ((p.dh + yo - 1) / yo), (frames + fpc - 1) / fpc);
    printf("level %d: %dx%d -> %dx%d, tile 128 x %d, grid %d x %d%s\n", l, p.sw, p.sh, p.dw, p.dh, yo, grids[l].x, grids[l].y,
           ok_level[l] ? "" : "  (does NOT fit the tile: skipped)");
  }
  CK(cudaFuncSetAttribute(k_resize_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  // the chain, both ways (level l reads level l - 1 of its own buffer)
  for (int l = 1; l < nl; l++) {
    dim3 g((w[l] + 127) / 128, h[l], frames);
    k_resize_ref<<<g, 128>>>(ref + off[l - 1], w[l - 1], h[l - 1], pitch[l - 1], fstride, ref + off[l], w[l], h[l], pitch[l],
                             fstride, d_tabs + xoff[l], d_tabs + yoff[l]);
    if (ok_level[l]) k_resize_tc<<<grids[l], kThreads, kSmem>>>(P[l]);
    else k_resize_ref<<<g, 128>>>(pyr + off[l - 1], w[l - 1], h[l - 1], pitch[l - 1], fstride, pyr + off[l], w[l], h[l],
                                  pitch[l], fstride, d_tabs + xoff[l], d_tabs + yoff[l]);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  {
    std::vector<uint8_t> a((size_t)fstride * frames), b((size_t)fstride * frames);
    CK(cudaMemcpy(a.data(), pyr, a.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), ref, b.size(), cudaMemcpyDeviceToHost));
    long long bad = 0, px = 0;
    for (int f = 0; f < frames; f++)
      for (int l = 1; l < nl; l++)
        for (int y = 0; y < h[l]; y++)
          for (int x = 0; x < w[l]; x++) {
            const size_t i = (size_t)f * fstride + off[l] + (size_t)y * pitch[l] + x;
            px++;
            if (a[i] != b[i]) {
              if (bad < 12) printf("  mismatch f %d level %d (%d, %d): got %d want %d\n", f, l, x, y, a[i], b[i]);
              bad++;
            }
          }
    printf("resize_tc: %lld of %lld pixels differ from the per-pixel kernel\n", bad, px);
    if (bad) return 1;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 20;
  auto chain = [&] {
    for (int l = 1; l < nl; l++)
      if (ok_level[l]) k_resize_tc<<<grids[l], kThreads, kSmem>>>(P[l]);
  };
  for (int i = 0; i < 3; i++) chain();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) chain();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  double pxs = 0;
  for (int l = 1; l < nl; l++) pxs += (double)w[l] * h[l];
  printf("resize_tc: %.3f ms per %d frames (levels 1..%d) = %.2f us / frame, %.2f Mpx / frame\n", ms, frames, nl - 1,
         1e3 * ms / frames, pxs / 1e6);
  return 0;
}
