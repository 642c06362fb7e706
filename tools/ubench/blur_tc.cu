// blur_tc — stand-alone bring-up / bench harness of the tensor-core form of cv::GaussianBlur(7x7, sigma 2, REFLECT_101)
// (src/ORBextractor.cc:1074-1076, OpenCV's fixed-point path: Q0.8 taps {18,34,48,56,48,34,18} on both axes, 16-bit row
// sums, one rounding (sum + 32768) >> 16). Checks itself against a plain per-pixel kernel and prints the rate.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o blur_tc blur_tc.cu -lcuda
//   timeout 120 ./blur_tc [frames=256] [w=640] [h=480] [levels=8] [frames per CTA=32] [pack 0|1] [variant 1|2]
//
// Idea. Both passes of the separable filter are banded matrix products over exact small integers, so they are
// u8 x u8 -> s32 GEMMs (tcgen05.mma kind::i8) and the ALU only moves bytes:
//   pass 1   D1[xo, r] = sum_k T[xo, k] * S[r, k]         T = 128 x 160 band of the taps (reflect-101 folded into the band
//                                                         of the tiles that touch the left / right edge), S = 128 source
//                                                         rows x 160 source columns, straight from TMA
//   split    D1 <= 255 * 256: hi / lo bytes of every sum -> two u8 planes H_hi, H_lo [xo][r] in shared memory
//   pass 2   D2hi[yo, xo] = sum_r W[yo, r] * H_hi[xo, r],  D2lo likewise;  W = 128 x 128 band over source rows, slot
//                                                         r = 126 carries the rounding constant (W = 128 x H_hi = 1)
//   out      byte 2 of (D2hi << 8) + D2lo, 128 consecutive x per thread -> swizzled tile -> TMA store
// Tile = 128 x 120 output pixels; all operands K-major in SWIZZLE_32B slabs of [128 rows x 32 bytes] (one slab = one
// K = 32 step). A CTA owns one tile position (level, tx, ty) — so its T and W are built once — and walks over frames
// with two ping-pong warpgroups (each: TMA -> MMA1 -> split -> MMA2 -> output), so that one group's tensor work runs
// under the other group's TMEM reads.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace {
constexpr int kTX = 128, kTY = 120;
constexpr int kSlab = 4096;
constexpr int kK1 = 5, kK2 = 4;
constexpr int kThreads = 256;
constexpr int oT = 0;
constexpr int oW = oT + kK1 * kSlab;
constexpr int oWG = oW + kK2 * kSlab;
constexpr int oS = 0;
constexpr int oH = oS + 2 * kK1 * kSlab;  // hi plane (kK2 slabs) then lo plane
constexpr int oOut = oH + 2 * kK2 * kSlab;
constexpr int kWGBytes = oOut + 128 * 128;
constexpr int oBar = oWG + 2 * kWGBytes;
constexpr int kSmem = oBar + 128 + 1024;
static_assert(kSmem <= 232448, "shared memory");
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // u8 x u8 -> s32

struct Params {
  CUtensorMap src[8];   // box 128 B x 128 rows, SWIZZLE_128B: source columns x0 - 16 .. x0 + 111
  CUtensorMap src2[8];  // box 32 B x 128 rows, SWIZZLE_32B: source columns x0 + 112 .. x0 + 143
  CUtensorMap dst[8];
  int w[8], h[8];
  int nlevels, frames, fpc;
};

__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  for (uint32_t spins = 0;; spins++) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(sptr(b)), "r"(parity) : "memory");
    if (ok) return;
    if (spins > (1u << 22)) {
      printf("blur_tc: barrier at %u of block (%d,%d) thread %d never completed (parity %u)\n", sptr(b), blockIdx.x,
             blockIdx.y, threadIdx.x, parity);
      __trap();
    }
  }
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
// K-major operand, SWIZZLE_32B: rows 32 B apart, 8-row groups 256 B apart (SBO = 16 x 16 B), version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc32(const void* p) {
  const uint64_t a = (sptr(p) & 0x3ffff) >> 4;
  return a | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}
// K-major operand inside a SWIZZLE_128B tile of 128-byte rows (8-row groups 1024 B apart); K step k starts 32 k bytes in
__device__ __forceinline__ uint64_t smem_desc128(const void* p) {
  const uint64_t a = (sptr(p) & 0x3ffff) >> 4;
  return a | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// pass-1 B operand of K step s: steps 0..3 live in the 128B-swizzled 16 KB tile, step 4 in the 32B-swizzled slab after it
__device__ __forceinline__ uint64_t s_desc(const uint8_t* stage, int s) {
  return s < 4 ? smem_desc128(stage + 32 * s) : smem_desc32(stage + 4 * kSlab);
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
#define TMEM_LD32_REGS(v)                                                                                              \
  "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),         \
      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),          \
      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),         \
      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {  // 32 columns
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : TMEM_LD32_REGS(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32_pack(uint32_t taddr, uint32_t (&v)[32]) {  // 64 columns, low halves, 2 / reg
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : TMEM_LD32_REGS(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wg_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }
__device__ __forceinline__ void st128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sptr(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// byte offset of element (row, k) of a K-major SWIZZLE_32B operand made of [128 x 32 B] slabs
__device__ __forceinline__ int sw32(int row, int k) {
  return (k >> 5) * kSlab + (row >> 3) * 256 + (row & 7) * 32 + ((((k >> 4) & 1) ^ ((row >> 2) & 1)) << 4) + (k & 15);
}
__device__ __forceinline__ int reflect101(int c, int n) { return c < 0 ? -c : (c >= n ? 2 * n - 2 - c : c); }

template <bool kPack>
__global__ void __launch_bounds__(kThreads, 1) k_blur_tc(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sT = smem + oT;
  uint8_t* sW = smem + oW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);  // per group: s_full[2], mma
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = tid >> 7, t = tid & 127;
  // tile position
  int c = blockIdx.x, l = 0, ntx = 0;
  for (;; l++) {
    if (l == P.nlevels) return;
    ntx = (P.w[l] + kTX - 1) / kTX;
    const int n = ntx * ((P.h[l] + kTY - 1) / kTY);
    if (c < n) break;
    c -= n;
  }
  const int w = P.w[l], h = P.h[l];
  const int ty = c / ntx, tx = c - ty * ntx;
  const int x0 = tx * kTX, y0 = ty * kTY;
  const int f0 = blockIdx.y * P.fpc, f1 = min(P.frames, f0 + P.fpc);

  if (tid == 0) {
    for (int i = 0; i < 6; i++) bar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- band matrices ----
  for (int i = tid; i < (kK1 + kK2) * kSlab / 16; i += kThreads) reinterpret_cast<uint4*>(sT)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    const int taps[7] = {18, 34, 48, 56, 48, 34, 18};
    if (g == 0) {
      const int x = x0 + t;
      if (x < w) {
#pragma unroll
        for (int j = 0; j < 7; j++) sT[sw32(t, reflect101(x + j - 3, w) - (x0 - 16))] += taps[j];
      }
    } else {
      const int y = y0 + t;
      if (t < kTY && y < h) {
#pragma unroll
        for (int j = 0; j < 7; j++) sW[sw32(t, reflect101(y + j - 3, h) - (y0 - 3))] += taps[j];
        sW[sw32(t, 126)] = 128;
      }
    }
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + g * 256;

  uint8_t* base = smem + oWG + g * kWGBytes;
  uint8_t* sS = base + oS;
  uint8_t* sHhi = base + oH;
  uint8_t* sHlo = sHhi + kK2 * kSlab;
  uint8_t* sOut = base + oOut;
  uint64_t* s_full = bars + 3 * g;
  uint64_t* mma_bar = bars + 3 * g + 2;
  const int quad = warp & 3;
  const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
  const int nf = (f1 - f0 - g + 1) / 2;  // frames f0 + g, f0 + g + 2, ...
  uint32_t mph = 0;

  auto load_S = [&](int stage, int f) {
    bar_expect(s_full + stage, kK1 * kSlab);
    tma_load3(&P.src[l], sS + stage * kK1 * kSlab, s_full + stage, x0 - 16, y0 - 3, f);
    tma_load3(&P.src2[l], sS + (stage * kK1 + 4) * kSlab, s_full + stage, x0 + 112, y0 - 3, f);
  };
  if (t == 0 && nf > 0) load_S(0, f0 + g);

  for (int i = 0; i < nf; i++) {
    const int st = i & 1, f = f0 + g + 2 * i;
    if (t == 0) {
      if (i + 1 < nf) load_S(st ^ 1, f + 2);
      bar_wait(s_full + st, (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < kK1; s++)
        mma_u8(tmem, smem_desc32(sT + s * kSlab), s_desc(sS + st * kK1 * kSlab, s), s > 0);
      mma_commit(mma_bar);
    }
    bar_wait(mma_bar, mph);
    mph ^= 1;
    tc_fence_after();
    // ---- split: D1 (lanes = xo, columns = source row r) -> H_hi / H_lo [xo][r] ----
    {
      const int rowoff = (t >> 3) * 256 + (t & 7) * 32, sw = (t >> 2) & 1;
      if (!kPack) {
#pragma unroll 1
        for (int s = 0; s < kK2; s++) {
          uint32_t v[32];
          tmem_ld32(lane_base + s * 32, v);
          tmem_ld_wait();
          if (s == kK2 - 1) {
            v[30] = 256;
            v[31] = 0;
          }
          uint32_t lo[8], hi[8];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const uint32_t a = __byte_perm(v[4 * q], v[4 * q + 1], 0x5410), b = __byte_perm(v[4 * q + 2], v[4 * q + 3], 0x5410);
            lo[q] = __byte_perm(a, b, 0x6420);
            hi[q] = __byte_perm(a, b, 0x7531);
          }
#pragma unroll
          for (int ch = 0; ch < 2; ch++) {
            const int off = s * kSlab + rowoff + ((ch ^ sw) << 4);
            st128(sHhi + off, hi[4 * ch], hi[4 * ch + 1], hi[4 * ch + 2], hi[4 * ch + 3]);
            st128(sHlo + off, lo[4 * ch], lo[4 * ch + 1], lo[4 * ch + 2], lo[4 * ch + 3]);
          }
        }
      } else {
#pragma unroll 1
        for (int s2 = 0; s2 < kK2 / 2; s2++) {  // 64 columns per load
          uint32_t v[32];
          tmem_ld32_pack(lane_base + s2 * 64, v);
          tmem_ld_wait();
          if (s2 == kK2 / 2 - 1) v[31] = 256;  // columns 126 (low half) = 256, 127 = 0
#pragma unroll
          for (int hs = 0; hs < 2; hs++) {
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
              lo[q] = __byte_perm(v[16 * hs + 2 * q], v[16 * hs + 2 * q + 1], 0x6420);
              hi[q] = __byte_perm(v[16 * hs + 2 * q], v[16 * hs + 2 * q + 1], 0x7531);
            }
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
              const int off = (2 * s2 + hs) * kSlab + rowoff + ((ch ^ sw) << 4);
              st128(sHhi + off, hi[4 * ch], hi[4 * ch + 1], hi[4 * ch + 2], hi[4 * ch + 3]);
              st128(sHlo + off, lo[4 * ch], lo[4 * ch + 1], lo[4 * ch + 2], lo[4 * ch + 3]);
            }
          }
        }
      }
    }
    tc_fence_before();
    proxy_fence();
    wg_sync(g);
    if (t == 0) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < kK2; s++) mma_u8(tmem + 128, smem_desc32(sW + s * kSlab), smem_desc32(sHhi + s * kSlab), s > 0);
#pragma unroll
      for (int s = 0; s < kK2; s++) mma_u8(tmem, smem_desc32(sW + s * kSlab), smem_desc32(sHlo + s * kSlab), s > 0);
      mma_commit(mma_bar);
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous tile's store has left sOut
    }
    bar_wait(mma_bar, mph);
    mph ^= 1;
    tc_fence_after();
    wg_sync(g);  // ... and everybody knows it
    // ---- output: lanes = yo, columns = xo ----
    {
      uint8_t* orow = sOut + t * 128;
      const int sw = t & 7;
      if (!kPack) {
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t hi[32], lo[32];
          tmem_ld32(lane_base + 128 + c0, hi);
          tmem_ld32(lane_base + c0, lo);
          tmem_ld_wait();
          uint32_t o[8];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            uint32_t v[4];
#pragma unroll
            for (int b = 0; b < 4; b++) v[b] = (hi[4 * q + b] << 8) + lo[4 * q + b];
            o[q] = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
          }
#pragma unroll
          for (int ch = 0; ch < 2; ch++)
            st128(orow + ((((c0 >> 4) + ch) ^ sw) << 4), o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 64) {
          uint32_t hi[32], lo[32];
          tmem_ld32_pack(lane_base + 128 + c0, hi);
          tmem_ld32_pack(lane_base + c0, lo);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const uint32_t a = hi[2 * q] + __byte_perm(lo[2 * q], 0, 0x4341);
            const uint32_t b = hi[2 * q + 1] + __byte_perm(lo[2 * q + 1], 0, 0x4341);
            o[q] = __byte_perm(a, b, 0x7531);
          }
#pragma unroll
          for (int ch = 0; ch < 4; ch++)
            st128(orow + ((((c0 >> 4) + ch) ^ sw) << 4), o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
        }
      }
    }
    tc_fence_before();
    proxy_fence();
    wg_sync(g);
    if (t == 0) tma_store3(&P.dst[l], sOut, x0, y0, f);
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_slot) : "memory");
}



// ---------------------------------------------------------------------------------------------------------------
// v4 = v1 with a lean issue path. ncu on v1: the issuing thread, not the tensor pipe, paced the kernel — inside
// `if (thread == 0)` ptxas cannot prove the descriptors / TMEM address uniform and wraps every UTCIMMA / UTMALDG in a
// uniformisation loop (R2UR.BROADCAST + BRA.U.ANY, ~16 instructions per MMA, ~170 clk per 64-clk MMA). Here the whole
// first warp of a group runs the issue path (warp index and TMEM base broadcast with shfl, so they are provably
// uniform), only the asynchronous instructions themselves sit under elect.sync, all descriptors are precomputed, and
// pass 2 is 4 MMAs of N = 256 (H_hi rows then H_lo rows per K step) instead of 8 of N = 128.
// ---------------------------------------------------------------------------------------------------------------
namespace v4 {
constexpr uint32_t kIdesc256 = (2u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_u8_256(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc256), "r"(accumulate), "r"(0u) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) k_blur_tc4(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sT = smem + oT;
  uint8_t* sW = smem + oW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);  // per group: s_full[2], mma
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int g = warp >> 2, t = tid & 127;
  int c = blockIdx.x, l = 0, ntx = 0;
  for (;; l++) {
    if (l == P.nlevels) return;
    ntx = (P.w[l] + kTX - 1) / kTX;
    const int n = ntx * ((P.h[l] + kTY - 1) / kTY);
    if (c < n) break;
    c -= n;
  }
  const int w = P.w[l], h = P.h[l];
  const int ty = c / ntx, tx = c - ty * ntx;
  const int x0 = tx * kTX, y0 = ty * kTY;
  const int f0 = blockIdx.y * P.fpc, f1 = min(P.frames, f0 + P.fpc);

  if (tid == 0) {
    for (int i = 0; i < 6; i++) bar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < (kK1 + kK2) * kSlab / 16; i += kThreads) reinterpret_cast<uint4*>(sT)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    const int taps[7] = {18, 34, 48, 56, 48, 34, 18};
    if (g == 0) {
      const int x = x0 + t;
      if (x < w) {
#pragma unroll
        for (int j = 0; j < 7; j++) sT[sw32(t, reflect101(x + j - 3, w) - (x0 - 16))] += taps[j];
      }
    } else {
      const int y = y0 + t;
      if (t < kTY && y < h) {
#pragma unroll
        for (int j = 0; j < 7; j++) sW[sw32(t, reflect101(y + j - 3, h) - (y0 - 3))] += taps[j];
        sW[sw32(t, 126)] = 128;
      }
    }
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tmem = tmem_base + g * 256;

  uint8_t* base = smem + oWG + g * kWGBytes;
  uint8_t* sS = base + oS;
  uint8_t* sH = base + oH;  // per K step: H_hi slab then H_lo slab (= one N = 256 operand)
  uint8_t* sOut = base + oOut;
  uint64_t* s_full = bars + 3 * g;
  uint64_t* mma_bar = bars + 3 * g + 2;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int nf = (f1 - f0 - g + 1) / 2;
  uint32_t mph = 0;
  const bool issuer = (warp & 3) == 0;

  // descriptors (issuer warp; uniform)
  uint64_t dT[kK1], dS[2][kK1], dW[kK2], dH[kK2];
#pragma unroll
  for (int s = 0; s < kK1; s++) {
    dT[s] = smem_desc32(sT + s * kSlab);
    dS[0][s] = s_desc(sS, s);
    dS[1][s] = s_desc(sS + kK1 * kSlab, s);
  }
#pragma unroll
  for (int s = 0; s < kK2; s++) {
    dW[s] = smem_desc32(sW + s * kSlab);
    dH[s] = smem_desc32(sH + 2 * s * kSlab);
  }

  auto load_S = [&](int stage, int f) {
    bar_expect(s_full + stage, kK1 * kSlab);
    tma_load3(&P.src[l], sS + stage * kK1 * kSlab, s_full + stage, x0 - 16, y0 - 3, f);
    tma_load3(&P.src2[l], sS + (stage * kK1 + 4) * kSlab, s_full + stage, x0 + 112, y0 - 3, f);
  };
  if (issuer && nf > 0) {
    if (elect_one()) load_S(0, f0 + g);
    __syncwarp();
  }

  for (int i = 0; i < nf; i++) {
    const int st = i & 1, f = f0 + g + 2 * i;
    if (issuer) {
      if (i + 1 < nf) {
        if (elect_one()) load_S(st ^ 1, f + 2);
        __syncwarp();
      }
      bar_wait(s_full + st, (i >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < kK1; s++) mma_u8(tmem, dT[s], st ? dS[1][s] : dS[0][s], s > 0);
        mma_commit(mma_bar);
      }
      __syncwarp();
    }
    bar_wait(mma_bar, mph);
    mph ^= 1;
    tc_fence_after();
    {
      const int rowoff = (t >> 3) * 256 + (t & 7) * 32, sw = (t >> 2) & 1;
#pragma unroll 1
      for (int s2 = 0; s2 < kK2 / 2; s2++) {
        uint32_t v[32];
        tmem_ld32_pack(lane_base + s2 * 64, v);
        tmem_ld_wait();
        if (s2 == kK2 / 2 - 1) v[31] = 256;
#pragma unroll
        for (int hs = 0; hs < 2; hs++) {
          uint32_t lo[8], hi[8];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            lo[q] = __byte_perm(v[16 * hs + 2 * q], v[16 * hs + 2 * q + 1], 0x6420);
            hi[q] = __byte_perm(v[16 * hs + 2 * q], v[16 * hs + 2 * q + 1], 0x7531);
          }
#pragma unroll
          for (int ch = 0; ch < 2; ch++) {
            const int off = 2 * (2 * s2 + hs) * kSlab + rowoff + ((ch ^ sw) << 4);
            st128(sH + off, hi[4 * ch], hi[4 * ch + 1], hi[4 * ch + 2], hi[4 * ch + 3]);
            st128(sH + kSlab + off, lo[4 * ch], lo[4 * ch + 1], lo[4 * ch + 2], lo[4 * ch + 3]);
          }
        }
      }
    }
    tc_fence_before();
    proxy_fence();
    wg_sync(g);
    if (issuer) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < kK2; s++) mma_u8_256(tmem, dW[s], dH[s], s > 0);  // columns 0..127 = hi, 128..255 = lo
        mma_commit(mma_bar);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    }
    bar_wait(mma_bar, mph);
    mph ^= 1;
    tc_fence_after();
    wg_sync(g);
    {
      uint8_t* orow = sOut + t * 128;
      const int sw = t & 7;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 64) {
        uint32_t hi[32], lo[32];
        tmem_ld32_pack(lane_base + c0, hi);
        tmem_ld32_pack(lane_base + 128 + c0, lo);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const uint32_t a = hi[2 * q] + __byte_perm(lo[2 * q], 0, 0x4341);
          const uint32_t b = hi[2 * q + 1] + __byte_perm(lo[2 * q + 1], 0, 0x4341);
          o[q] = __byte_perm(a, b, 0x7531);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ch++)
          st128(orow + ((((c0 >> 4) + ch) ^ sw) << 4), o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
      }
    }
    tc_fence_before();
    proxy_fence();
    wg_sync(g);
    if (issuer) {
      if (elect_one()) tma_store3(&P.dst[l], sOut, x0, y0, f);
      __syncwarp();
    }
  }
  // the elected lane may differ from call to call only in theory (elect.sync picks the lowest active lane); bulk groups
  // are per thread, so the final wait is done by every lane of the issuer warp
  if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}
}  // namespace v4

// ---------------------------------------------------------------------------------------------------------------
// v2: warp-specialised pipeline. warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = split (D1 -> H), warps 6..9 =
// output (D2 -> tile -> TMA store). TMEM: D1 double-buffered (2 x 128 columns), D2 as two independently recycled
// halves of 64 output columns (each 64 hi + 64 lo columns). Shared memory: S 3 stages, H 2 stages, output tile 2 stages.
// ---------------------------------------------------------------------------------------------------------------
namespace v2 {
constexpr int kThreads2 = 320;
constexpr int kSStages = 3;
constexpr int oT2 = 0;
constexpr int oW2 = oT2 + kK1 * kSlab;
constexpr int oS2 = oW2 + kK2 * kSlab;
constexpr int oH2 = oS2 + kSStages * kK1 * kSlab;   // per stage: hi plane (kK2 slabs) then lo plane
constexpr int oOut2 = oH2 + 2 * 2 * kK2 * kSlab;
constexpr int oBar2 = oOut2 + 2 * 128 * 128;
constexpr int kSmem2 = oBar2 + 256 + 1024;
static_assert(kSmem2 <= 232448, "shared memory");
constexpr uint32_t kIdesc64 = (2u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void mma_u8_n64(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc64), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(b)) : "memory");
}

__global__ void __launch_bounds__(kThreads2, 1) k_blur_tc2(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sT = smem + oT2;
  uint8_t* sW = smem + oW2;
  uint8_t* sS = smem + oS2;
  uint8_t* sH = smem + oH2;
  uint8_t* sOut = smem + oOut2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar2);
  uint64_t *s_full = bars, *s_empty = bars + 3, *d1_full = bars + 6, *d1_empty = bars + 8, *h_full = bars + 10,
           *h_empty = bars + 12, *d2_full = bars + 14, *d2_empty = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int c = blockIdx.x, l = 0, ntx = 0;
  for (;; l++) {
    if (l == P.nlevels) return;
    ntx = (P.w[l] + kTX - 1) / kTX;
    const int n = ntx * ((P.h[l] + kTY - 1) / kTY);
    if (c < n) break;
    c -= n;
  }
  const int w = P.w[l], h = P.h[l];
  const int ty = c / ntx, tx = c - ty * ntx;
  const int x0 = tx * kTX, y0 = ty * kTY;
  const int f0 = blockIdx.y * P.fpc, f1 = min(P.frames, f0 + P.fpc);
  const int n = f1 - f0;

  if (tid == 0) {
    for (int i = 0; i < 3; i++) {
      bar_init(s_full + i, 1);
      bar_init(s_empty + i, 1);
    }
    for (int i = 0; i < 2; i++) {
      bar_init(d1_full + i, 1);
      bar_init(d1_empty + i, 128);
      bar_init(h_full + i, 128);
      bar_init(h_empty + i, 1);
      bar_init(d2_full + i, 1);
      bar_init(d2_empty + i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < (kK1 + kK2) * kSlab / 16; i += kThreads2) reinterpret_cast<uint4*>(sT)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    const int taps[7] = {18, 34, 48, 56, 48, 34, 18};
    if (tid < 128) {
      const int x = x0 + tid;
      if (x < w) {
#pragma unroll
        for (int j = 0; j < 7; j++) sT[sw32(tid, reflect101(x + j - 3, w) - (x0 - 16))] += taps[j];
      }
    } else if (tid < 256) {
      const int t = tid - 128, y = y0 + t;
      if (t < kTY && y < h) {
#pragma unroll
        for (int j = 0; j < 7; j++) sW[sw32(t, reflect101(y + j - 3, h) - (y0 - 3))] += taps[j];
        sW[sw32(t, 126)] = 128;
      }
    }
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n; i++) {
        const int st = i % kSStages;
        bar_wait(s_empty + st, ((i / kSStages) & 1) ^ 1);
        bar_expect(s_full + st, kK1 * kSlab);
        tma_load3(&P.src[l], sS + st * kK1 * kSlab, s_full + st, x0 - 16, y0 - 3, f0 + i);
        tma_load3(&P.src2[l], sS + (st * kK1 + 4) * kSlab, s_full + st, x0 + 112, y0 - 3, f0 + i);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      auto mma1 = [&](int j) {
        const int st = j % kSStages, b = j & 1;
        bar_wait(s_full + st, (j / kSStages) & 1);
        bar_wait(d1_empty + b, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < kK1; s++)
          mma_u8(tmem + b * 128, smem_desc32(sT + s * kSlab), s_desc(sS + st * kK1 * kSlab, s), s > 0);
        mma_commit(s_empty + st);
        mma_commit(d1_full + b);
      };
      auto mma2 = [&](int j) {
        const int b = j & 1;
        uint8_t* hi = sH + b * 2 * kK2 * kSlab;
        uint8_t* lo = hi + kK2 * kSlab;
        bar_wait(h_full + b, (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          bar_wait(d2_empty + hf, (j & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem + 256 + hf * 128;
#pragma unroll
          for (int s = 0; s < kK2; s++)
            mma_u8_n64(d, smem_desc32(sW + s * kSlab), smem_desc32(hi + s * kSlab + hf * 2048), s > 0);
#pragma unroll
          for (int s = 0; s < kK2; s++)
            mma_u8_n64(d + 64, smem_desc32(sW + s * kSlab), smem_desc32(lo + s * kSlab + hf * 2048), s > 0);
          mma_commit(d2_full + hf);
        }
        mma_commit(h_empty + b);
      };
      if (n > 0) mma1(0);
      for (int i = 0; i < n; i++) {
        if (i + 1 < n) mma1(i + 1);
        mma2(i);
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ---- split: D1 (lanes = xo, columns = source row r) -> H_hi / H_lo [xo][r] ----
    const int t = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int rowoff = (t >> 3) * 256 + (t & 7) * 32, sw = (t >> 2) & 1;
    for (int i = 0; i < n; i++) {
      const int b = i & 1;
      uint8_t* sHhi = sH + b * 2 * kK2 * kSlab;
      uint8_t* sHlo = sHhi + kK2 * kSlab;
      bar_wait(d1_full + b, (i >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32_pack(lane_base + b * 128, v0);
      tmem_ld32_pack(lane_base + b * 128 + 64, v1);
      tmem_ld_wait();
      tc_fence_before();
      bar_arrive(d1_empty + b);
      v1[31] = 256;  // column 126 (low half) = 256 -> H_hi = 1: the rounding slot; column 127 = 0
      bar_wait(h_empty + b, ((i >> 1) & 1) ^ 1);
#pragma unroll
      for (int s = 0; s < kK2; s++) {
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const uint32_t a = s < 2 ? v0[16 * s + 2 * q] : v1[16 * (s - 2) + 2 * q];
          const uint32_t bb = s < 2 ? v0[16 * s + 2 * q + 1] : v1[16 * (s - 2) + 2 * q + 1];
          lo[q] = __byte_perm(a, bb, 0x6420);
          hi[q] = __byte_perm(a, bb, 0x7531);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          const int off = s * kSlab + rowoff + ((ch ^ sw) << 4);
          st128(sHhi + off, hi[4 * ch], hi[4 * ch + 1], hi[4 * ch + 2], hi[4 * ch + 3]);
          st128(sHlo + off, lo[4 * ch], lo[4 * ch + 1], lo[4 * ch + 2], lo[4 * ch + 3]);
        }
      }
      proxy_fence();
      bar_arrive(h_full + b);
    }
  } else {
    // ---- output: D2 halves (lanes = yo, columns = 64 xo: hi then lo) -> swizzled tile -> TMA store ----
    const int t = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    const int sw = t & 7;
    const bool leader = warp == 6 && lane == 0;
    for (int i = 0; i < n; i++) {
      uint8_t* orow = sOut + (i & 1) * 128 * 128 + t * 128;
      if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store of tile i - 2 left the buffer
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        bar_wait(d2_full + hf, i & 1);
        tc_fence_after();
        uint32_t hi[32], lo[32];
        tmem_ld32_pack(lane_base + hf * 128, hi);
        tmem_ld32_pack(lane_base + hf * 128 + 64, lo);
        tmem_ld_wait();
        tc_fence_before();
        bar_arrive(d2_empty + hf);
        uint32_t o[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const uint32_t a = hi[2 * q] + __byte_perm(lo[2 * q], 0, 0x4341);
          const uint32_t bb = hi[2 * q + 1] + __byte_perm(lo[2 * q + 1], 0, 0x4341);
          o[q] = __byte_perm(a, bb, 0x7531);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ch++)
          st128(orow + (((hf * 4 + ch) ^ sw) << 4), o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
      }
      proxy_fence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) tma_store3(&P.dst[l], sOut + (i & 1) * 128 * 128, x0, y0, f0 + i);
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
}  // namespace v2



// ---------------------------------------------------------------------------------------------------------------
// v3: the row sums never touch shared memory. Pass 2 takes A = H from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc):
// D2[xo, yo] = sum_r H[xo, r] * W[yo, r], lanes = xo — the layout D1 already has — so the split warps repack D1 into
// bytes with tcgen05.st (4 source rows per 32-bit cell, 8 cells per K = 32 step) and the B operand is the constant W.
// Shared-memory traffic per tile drops from ~188 KB to ~124 KB (no H planes written / read, no W re-read as A). The price
// is a transposed accumulator (a thread owns a COLUMN of the output): 4x4 byte transposes with two shuffles per word.
// warp 0 = TMA, warp 1 = MMA, warps 2..5 = split, warps 6..9 = output.
// TMEM columns: D1 0..127, D2hi 128..255, D2lo 256..383, H[2] at 384 + 64 b (hi 32 cells, lo 32 cells).
// ---------------------------------------------------------------------------------------------------------------
namespace v3 {
constexpr int kThreads3 = 320;
constexpr int kSStages = 3;
constexpr int oT3 = 0;
constexpr int oW3 = oT3 + kK1 * kSlab;
constexpr int oS3 = oW3 + kK2 * kSlab;
constexpr int oOut3 = oS3 + kSStages * kK1 * kSlab;
constexpr int oBar3 = oOut3 + 2 * 128 * 128;
constexpr int kSmem3 = oBar3 + 256 + 1024;

__device__ __forceinline__ void mma_u8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(kIdesc), "r"(accumulate), "r"(0u) : "memory");
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
__device__ __forceinline__ void bar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(b)) : "memory");
}

__global__ void __launch_bounds__(kThreads3, 1) k_blur_tc3(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sT = smem + oT3;
  uint8_t* sW = smem + oW3;
  uint8_t* sS = smem + oS3;
  uint8_t* sOut = smem + oOut3;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar3);
  uint64_t *s_full = bars, *s_empty = bars + 3, *d1_full = bars + 6, *d1_empty = bars + 7, *h_full = bars + 8,
           *h_empty = bars + 10, *d2_full = bars + 12, *d2_empty = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  int c = blockIdx.x, l = 0, ntx = 0;
  for (;; l++) {
    if (l == P.nlevels) return;
    ntx = (P.w[l] + kTX - 1) / kTX;
    const int n = ntx * ((P.h[l] + kTY - 1) / kTY);
    if (c < n) break;
    c -= n;
  }
  const int w = P.w[l], h = P.h[l];
  const int ty = c / ntx, tx = c - ty * ntx;
  const int x0 = tx * kTX, y0 = ty * kTY;
  const int f0 = blockIdx.y * P.fpc, f1 = min(P.frames, f0 + P.fpc);
  const int n = f1 - f0;

  if (tid == 0) {
    for (int i = 0; i < 3; i++) {
      bar_init(s_full + i, 1);
      bar_init(s_empty + i, 1);
    }
    bar_init(d1_full, 1);
    bar_init(d1_empty, 128);
    bar_init(d2_full, 1);
    bar_init(d2_empty, 128);
    for (int i = 0; i < 2; i++) {
      bar_init(h_full + i, 128);
      bar_init(h_empty + i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < (kK1 + kK2) * kSlab / 16; i += kThreads3) reinterpret_cast<uint4*>(sT)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    const int taps[7] = {18, 34, 48, 56, 48, 34, 18};
    if (tid < 128) {
      const int x = x0 + tid;
      if (x < w) {
#pragma unroll
        for (int j = 0; j < 7; j++) sT[sw32(tid, reflect101(x + j - 3, w) - (x0 - 16))] += taps[j];
      }
    } else if (tid < 256) {
      const int t = tid - 128, y = y0 + t;
      if (t < kTY && y < h) {
#pragma unroll
        for (int j = 0; j < 7; j++) sW[sw32(t, reflect101(y + j - 3, h) - (y0 - 3))] += taps[j];
        sW[sw32(t, 126)] = 128;
      }
    }
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    for (int i = 0; i < n; i++) {
      const int st = i % kSStages;
      bar_wait(s_empty + st, ((i / kSStages) & 1) ^ 1);
      if (v4::elect_one()) {
        bar_expect(s_full + st, kK1 * kSlab);
        tma_load3(&P.src[l], sS + st * kK1 * kSlab, s_full + st, x0 - 16, y0 - 3, f0 + i);
        tma_load3(&P.src2[l], sS + (st * kK1 + 4) * kSlab, s_full + st, x0 + 112, y0 - 3, f0 + i);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // the whole warp runs the issue path (uniform descriptors); only the asynchronous instructions sit under elect.sync
    uint64_t dT[kK1], dS[kSStages][kK1], dW[kK2];
#pragma unroll
    for (int s = 0; s < kK1; s++) {
      dT[s] = smem_desc32(sT + s * kSlab);
#pragma unroll
      for (int st = 0; st < kSStages; st++) dS[st][s] = s_desc(sS + st * kK1 * kSlab, s);
    }
#pragma unroll
    for (int s = 0; s < kK2; s++) dW[s] = smem_desc32(sW + s * kSlab);
    auto mma1 = [&](int j) {
      const int st = j % kSStages;
      bar_wait(s_full + st, (j / kSStages) & 1);
      bar_wait(d1_empty, (j & 1) ^ 1);
      tc_fence_after();
      if (v4::elect_one()) {
#pragma unroll
        for (int s = 0; s < kK1; s++) mma_u8(tmem, dT[s], st == 0 ? dS[0][s] : (st == 1 ? dS[1][s] : dS[2][s]), s > 0);
        mma_commit(s_empty + st);
        mma_commit(d1_full);
      }
      __syncwarp();
    };
    auto mma2 = [&](int j) {
      const int b = j & 1;
      bar_wait(h_full + b, (j >> 1) & 1);
      bar_wait(d2_empty, (j & 1) ^ 1);
      tc_fence_after();
      const uint32_t hcol = tmem + 384 + 64 * b;
      if (v4::elect_one()) {
#pragma unroll
        for (int s = 0; s < kK2; s++) mma_u8_ts(tmem + 128, hcol + 8 * s, dW[s], s > 0);
#pragma unroll
        for (int s = 0; s < kK2; s++) mma_u8_ts(tmem + 256, hcol + 32 + 8 * s, dW[s], s > 0);
        mma_commit(h_empty + b);
        mma_commit(d2_full);
      }
      __syncwarp();
    };
    if (n > 0) mma1(0);
    for (int i = 0; i < n; i++) {
      if (i + 1 < n) mma1(i + 1);
      mma2(i);
    }
  } else if (warp < 6) {
    // ---- split: D1 (lanes = xo, columns = source row r) -> packed bytes H_hi / H_lo in tensor memory ----
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < n; i++) {
      const int b = i & 1;
      bar_wait(d1_full, i & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32_pack(lane_base, v0);
      tmem_ld32_pack(lane_base + 64, v1);
      tmem_ld_wait();
      tc_fence_before();
      bar_arrive(d1_empty);
      v1[31] = 256;  // column 126 = 256 -> H_hi = 1: the rounding slot; column 127 = 0
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int q = 0; q < 16; q++) {
        lo[q] = __byte_perm(v0[2 * q], v0[2 * q + 1], 0x6420);
        hi[q] = __byte_perm(v0[2 * q], v0[2 * q + 1], 0x7531);
        lo[16 + q] = __byte_perm(v1[2 * q], v1[2 * q + 1], 0x6420);
        hi[16 + q] = __byte_perm(v1[2 * q], v1[2 * q + 1], 0x7531);
      }
      bar_wait(h_empty + b, ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      tmem_st32(lane_base + 384 + 64 * b, hi);
      tmem_st32(lane_base + 384 + 64 * b + 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      bar_arrive(h_full + b);
    }
  } else {
    // ---- output: D2 (lanes = xo, columns = yo) -> 4x4 byte transposes inside lane quads -> swizzled tile [yo][x] ----
    const int quad = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    const bool leader = warp == 6 && lane == 0;
    const uint32_t sel1 = (lane & 1) ? 0x3715u : 0x6240u, sel2 = (lane & 2) ? 0x3276u : 0x5410u;
    const int xw = quad * 32 + 4 * (lane >> 2);  // first of the 4 x this lane ends up with
    const int ch = xw >> 4, inner = xw & 15;
    for (int i = 0; i < n; i++) {
      uint8_t* obuf = sOut + (i & 1) * 128 * 128;
      if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store of tile i - 2 left the buffer
      asm volatile("bar.sync 1, 128;" ::: "memory");
      bar_wait(d2_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 64) {
        uint32_t hi[32], lo[32];
        tmem_ld32_pack(lane_base + 128 + c0, hi);
        tmem_ld32_pack(lane_base + 256 + c0, lo);
        tmem_ld_wait();
        if (c0 == 64) {
          tc_fence_before();
          bar_arrive(d2_empty);
        }
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const uint32_t a = hi[2 * q] + __byte_perm(lo[2 * q], 0, 0x4341);
          const uint32_t bb = hi[2 * q + 1] + __byte_perm(lo[2 * q + 1], 0, 0x4341);
          uint32_t wv = __byte_perm(a, bb, 0x7531);  // this x, yo = c0 + 4 q .. + 3
          const uint32_t t1 = __shfl_xor_sync(0xffffffffu, wv, 1);
          wv = __byte_perm(wv, t1, sel1);
          const uint32_t t2 = __shfl_xor_sync(0xffffffffu, wv, 2);
          wv = __byte_perm(wv, t2, sel2);            // yo = c0 + 4 q + (lane & 3), x = xw .. xw + 3
          const int yo = c0 + 4 * q + (lane & 3);
          *reinterpret_cast<uint32_t*>(obuf + yo * 128 + ((ch ^ (yo & 7)) << 4) + inner) = wv;
        }
      }
      proxy_fence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) tma_store3(&P.dst[l], obuf, x0, y0, f0 + i);
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
}  // namespace v3

// ---------------------------------------------------------------------------------------------------------------
// TMEM read-rate probe: every warp of a CTA reads its lane quadrant 32 columns at a time; prints bytes / clk / SM.
// ---------------------------------------------------------------------------------------------------------------
template <bool kPk, int kInFlight>
__global__ void __launch_bounds__(512, 1) k_ldtm_probe(int iters, long long* clocks, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    uint32_t v[kInFlight][32];
#pragma unroll
    for (int k = 0; k < kInFlight; k++) {
      const uint32_t col = ((i * kInFlight + k) * (kPk ? 64 : 32) + (warp >> 2) * 64) & 511 & ~(kPk ? 63u : 31u);
      if (kPk) tmem_ld32_pack(base + col, v[k]);
      else tmem_ld32(base + col, v[k]);
    }
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < kInFlight; k++) acc += v[k][0] ^ v[k][31];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <bool kPk, int kInFlight>
static void run_probe(int warps) {
  long long* d;
  uint32_t* sink;
  cudaMalloc(&d, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  k_ldtm_probe<kPk, kInFlight><<<148, warps * 32>>>(iters, d, sink);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("probe failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return;
  }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double clk = (double)h[0] / iters;
  const double cells = (double)warps * kInFlight * 32 * (kPk ? 64 : 32);  // 32-bit TMEM cells touched per iteration
  printf("ldtm probe: %2d warps, %d loads in flight, pack %d: %.1f clk / iteration, %.1f TMEM B / clk / SM, %.1f register B / clk / SM\n",
         warps, kInFlight, (int)kPk, clk, cells * 4 / clk, (double)warps * kInFlight * 32 * 32 * 4 / clk);
  cudaFree(d);
  cudaFree(sink);
}

// ---------------------------------------------------------------------------------------------------------------
// MMA rate probe: one thread issues `reps` back-to-back tcgen05.mma kind::i8 (M = 128, K = 32) on resident operands and
// waits for the commit; prints clocks per MMA. mode 0: A, B in SWIZZLE_32B slabs; 1: A, B in a SWIZZLE_128B tile;
// 2: A from tensor memory, B SWIZZLE_32B. N = 64 / 128 / 256.
// ---------------------------------------------------------------------------------------------------------------
template <int kN, int kMode>
__global__ void __launch_bounds__(128, 1) k_mma_probe(int reps, long long* clocks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    bar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sptr(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t idesc = (2u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < reps; i++) {
      const int s = i & 3;  // 4 K steps, like a real tile
      uint64_t da, db;
      if (kMode == 1) {
        da = smem_desc128(smem + 32 * s);
        db = smem_desc128(smem + 32 * 1024 + 32 * s);
      } else {
        da = smem_desc32(smem + s * kSlab);
        db = smem_desc32(smem + 32 * 1024 + s * (kN * 32));
      }
      if (kMode == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
            ::"r"(tmem), "r"(tmem + 384 + 8 * s), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
            ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
      }
    }
    mma_commit(bar);
    bar_wait(bar, 0);
    const long long t1 = clock64();
    clocks[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int kN, int kMode>
static void run_mma_probe(int ctas) {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  const int reps = 4096, smem = 96 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(k_mma_probe<kN, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_mma_probe<kN, kMode><<<ctas, 128, smem>>>(reps, d);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("mma probe failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return;
  }
  long long h[148];
  cudaMemcpy(h, d, ctas * 8, cudaMemcpyDeviceToHost);
  const char* names[3] = {"A, B SWIZZLE_32B slabs", "A, B in a SWIZZLE_128B tile", "A in TMEM, B SWIZZLE_32B"};
  printf("mma probe: M 128 N %3d K 32 u8, %-28s %3d CTAs: %.1f clk / MMA (ideal %d)\n", kN, names[kMode], ctas,
         (double)h[0] / reps, kN / 2);
  cudaFree(d);
}

// plain reference: one thread per pixel
__global__ void k_blur_ref(const uint8_t* src, uint8_t* dst, int w, int h, int pitch, int64_t fstride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (x >= w) return;
  const int taps[7] = {18, 34, 48, 56, 48, 34, 18};
  const uint8_t* s = src + f * fstride;
  uint32_t acc = 32768;
  for (int i = 0; i < 7; i++) {
    const int yy = reflect101(y + i - 3, h);
    uint32_t r = 0;
    for (int j = 0; j < 7; j++) r += taps[j] * s[(int64_t)yy * pitch + reflect101(x + j - 3, w)];
    acc += taps[i] * r;
  }
  dst[f * fstride + (int64_t)y * pitch + x] = (uint8_t)(acc >> 16);
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
  if (argc > 1 && !strcmp(argv[1], "mmaprobe")) {
    run_mma_probe<128, 0>(148);
    run_mma_probe<128, 1>(148);
    run_mma_probe<128, 2>(148);
    run_mma_probe<64, 0>(148);
    run_mma_probe<64, 2>(148);
    run_mma_probe<256, 0>(148);
    run_mma_probe<256, 1>(148);
    run_mma_probe<256, 2>(148);
    run_mma_probe<128, 0>(1);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "probe")) {
    run_probe<false, 1>(4);
    run_probe<false, 2>(4);
    run_probe<false, 4>(4);
    run_probe<false, 2>(8);
    run_probe<false, 2>(16);
    run_probe<true, 1>(4);
    run_probe<true, 2>(4);
    run_probe<true, 2>(8);
    run_probe<false, 2>(1);
    return 0;
  }
  const int frames = argc > 1 ? atoi(argv[1]) : 256;
  const int W = argc > 2 ? atoi(argv[2]) : 640, H = argc > 3 ? atoi(argv[3]) : 480;
  const int nl = argc > 4 ? atoi(argv[4]) : 8;
  const int fpc = argc > 5 ? atoi(argv[5]) : 32;
  const int pack = argc > 6 ? atoi(argv[6]) : 0;
  const int variant = argc > 7 ? atoi(argv[7]) : 1;
  Params P;
  memset(&P, 0, sizeof(P));
  P.nlevels = nl;
  P.frames = frames;
  P.fpc = fpc;
  int pitch[8];
  int64_t off[8], total = 0;
  float sf = 1.f;
  int classes = 0;
  for (int l = 0; l < nl; l++) {
    P.w[l] = (int)lrintf((float)W / sf);
    P.h[l] = (int)lrintf((float)H / sf);
    sf *= 1.2f;
    pitch[l] = (P.w[l] + 63) / 64 * 64;
    off[l] = total;
    total += (int64_t)pitch[l] * P.h[l];
    classes += ((P.w[l] + kTX - 1) / kTX) * ((P.h[l] + kTY - 1) / kTY);
  }
  const int64_t fstride = (total + 255) / 256 * 256;
  uint8_t *src, *dst, *ref;
  CK(cudaMalloc(&src, fstride * frames));
  CK(cudaMalloc(&dst, fstride * frames));
  CK(cudaMalloc(&ref, fstride * frames));
  {
    std::vector<uint8_t> hsrc((size_t)fstride * frames);
    uint32_t s = 12345;
    for (size_t i = 0; i < hsrc.size(); i++) {
      s = s * 1664525u + 1013904223u;
      // mixture: smooth-ish ramps and extremes so that saturated sums (255 * 256) occur
      const uint32_t r = s >> 24;
      hsrc[i] = (i / 977) % 7 == 0 ? 255 : ((i / 1013) % 11 == 0 ? 0 : (uint8_t)r);
    }
    CK(cudaMemcpy(src, hsrc.data(), hsrc.size(), cudaMemcpyHostToDevice));
  }
  CK(cudaMemset(dst, 0xcd, fstride * frames));
  CK(cudaMemset(ref, 0xcd, fstride * frames));

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  for (int l = 0; l < nl; l++) {
    const cuuint64_t dims[3] = {(cuuint64_t)P.w[l], (cuuint64_t)P.h[l], (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch[l], (cuuint64_t)fstride};
    const cuuint32_t es[3] = {1, 1, 1};
    const cuuint32_t box_s[3] = {128, 128, 1}, box_s2[3] = {32, 128, 1}, box_d[3] = {128, (cuuint32_t)kTY, 1};
    CUresult r1 = enc(&P.src[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, src + off[l], dims, strides, box_s, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r3 = enc(&P.src2[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, src + off[l], dims, strides, box_s2, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r3 != CUDA_SUCCESS) r1 = r3;
    CUresult r2 = enc(&P.dst[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dst + off[l], dims, strides, box_d, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      printf("tensor map encode failed: level %d (%d x %d) -> %d %d\n", l, P.w[l], P.h[l], (int)r1, (int)r2);
      return 2;
    }
  }
  for (int l = 0; l < nl; l++) {
    dim3 grid((P.w[l] + 127) / 128, P.h[l], frames);
    k_blur_ref<<<grid, 128>>>(src + off[l], ref + off[l], P.w[l], P.h[l], pitch[l], fstride);
  }
  CK(cudaDeviceSynchronize());

  auto kern = variant == 4 ? v4::k_blur_tc4 : variant == 3 ? v3::k_blur_tc3 : variant == 2 ? v2::k_blur_tc2 : (pack ? k_blur_tc<true> : k_blur_tc<false>);
  const int kSmemUsed = variant == 4 ? kSmem : variant == 3 ? v3::kSmem3 : variant == 2 ? v2::kSmem2 : kSmem;
  const int kThreadsUsed = (variant == 2 || variant == 3) ? 320 : kThreads;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemUsed));
  dim3 grid(classes, (frames + fpc - 1) / fpc);
  printf("blur_tc: %d frames of %dx%d x %d levels, %d tile positions, grid %d x %d, %d B smem, pack %d, variant %d\n",
         frames, W, H, nl, classes, grid.x, grid.y, kSmemUsed, pack, variant);
  kern<<<grid, kThreadsUsed, kSmemUsed>>>(P);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  // compare
  {
    std::vector<uint8_t> a((size_t)fstride * frames), b((size_t)fstride * frames);
    CK(cudaMemcpy(a.data(), dst, a.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), ref, b.size(), cudaMemcpyDeviceToHost));
    long long bad = 0, px = 0;
    for (int f = 0; f < frames; f++)
      for (int l = 0; l < nl; l++)
        for (int y = 0; y < P.h[l]; y++)
          for (int x = 0; x < P.w[l]; x++) {
            const size_t i = (size_t)f * fstride + off[l] + (size_t)y * pitch[l] + x;
            px++;
            if (a[i] != b[i]) {
              if (bad < 12) printf("  mismatch f %d level %d (%d, %d): got %d want %d\n", f, l, x, y, a[i], b[i]);
              bad++;
            }
          }
    // a TMA store clips the box at the tensor bound in whole 16-byte units (measured: bytes w .. round_up(w, 16) - 1 of
    // every row are written); the padding beyond that must be untouched
    long long pad_bad = 0;
    for (int l = 0; l < nl; l++)
      for (int y = 0; y < P.h[l]; y++)
        for (int x = (P.w[l] + 15) / 16 * 16; x < pitch[l]; x++)
          if (a[off[l] + (size_t)y * pitch[l] + x] != 0xcd) pad_bad++;
    printf("blur_tc: %lld of %lld pixels differ from the per-pixel kernel; %lld padding bytes touched\n", bad, px, pad_bad);
    if (bad || pad_bad) return 1;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 20;
  for (int i = 0; i < 3; i++) kern<<<grid, kThreadsUsed, kSmemUsed>>>(P);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) kern<<<grid, kThreadsUsed, kSmemUsed>>>(P);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  double pxs = 0;
  for (int l = 0; l < nl; l++) pxs += (double)P.w[l] * P.h[l];
  printf("blur_tc: %.3f ms per %d frames = %.2f us / frame, %.0f GB/s of 2 x pixels (%.2f Mpx / frame)\n", ms, frames,
         1e3 * ms / frames, 2 * pxs * frames / ms / 1e6, pxs / 1e6);
  return 0;
}
