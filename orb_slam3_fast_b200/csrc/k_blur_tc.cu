// k_blur_tc.cu — tensor-core form of cv::GaussianBlur(level, 7x7, sigma 2, REFLECT_101) (src/ORBextractor.cc:1074-1076;
// OpenCV's fixed-point path: Q0.8 taps {18,34,48,56,48,34,18} on both axes, exact 16-bit row sums, ONE rounding
// (sum + 32768) >> 16). k_blur7 in k_describe.cu stays as the fallback (caller-owned level 0 that TMA cannot address,
// more than 8 levels, ORBX_BLUR_TC=0).
//
// Both passes of the separable filter are banded matrix products over small exact integers, i.e. u8 x u8 -> s32 GEMMs
// (tcgen05.mma kind::i8) — the CUDA cores only move bytes (k_blur7 needs ~20 instructions per pixel and is issue bound at
// 30 % of the HBM roof):
//   pass 1   D1[xo, r]   = sum_k T[xo, k] * S[r, k]      A = T: 128 x 160 band of the taps over source columns x0-16 .. x0+143
//                                                        (reflect-101 folded into the band of the tiles that touch the
//                                                        left / right edge; columns outside the level arrive as TMA
//                                                        zeros and carry weight 0); B = S: 128 source rows, straight from TMA
//   split    D1 <= 255 * 256: the hi / lo bytes of every row sum, repacked 4 source rows per 32-bit cell and written back
//            to TENSOR MEMORY (tcgen05.st) — the row sums never touch shared memory
//   pass 2   D2hi[xo, yo] = sum_r H_hi[xo, r] * W[yo, r], D2lo likewise: A = H from tensor memory (lanes = xo, the
//                                                        layout D1 already has), B = W: 128 x 128 band over source rows
//                                                        y0-3 .. y0+124 (reflect-101 folded in for the top / bottom tiles);
//                                                        slot r = 126 carries the rounding: W = 128, H_hi = 1 -> + 32768
//   out      byte 2 of (D2hi << 8) + D2lo. A thread owns a COLUMN of the tile (lane = xo), so words of 4 consecutive yo are
//            transposed 4 x 4 inside lane quads (2 shuffles + 2 PRMT per word) -> 128B-swizzled tile [yo][x] -> TMA store
//            (clipped at the level's bounds in whole 16-byte units: bytes w .. round_up(w, 16) - 1 of a row are written;
//            they are row padding — pitch = round_up(w, 64))
// Everything is exact integer arithmetic: D1 < 2^16, D2hi, D2lo < 2^16 + 2^15, so the 16-bit packed TMEM loads
// (tcgen05.ld ... .pack::16b) lose nothing.
//
// Tile = 128 x 120 output pixels; a CTA owns one tile position (level, tx, ty) — its T and W are built once — and walks
// over `fpc` frames. warp 0 = TMA producer (3 stages), warp 1 = MMA issuer, warps 2..5 = split, warps 6..9 = output.
// TMEM columns: D1 0..127, D2hi 128..255, D2lo 256..383, H[2] at 384 + 64 b (hi 32 cells, lo 32 cells).
// Shared memory: 128 KB (T 20, W 16, S 3 x 20, output tiles 2 x 16) — other kernels' blocks still fit beside it.
//
// Measured on B200 with the stand-alone form (tools/ubench/blur_tc.cu, variants 1-4; 0 of 9.7e8 pixels differ from a
// per-pixel kernel in every variant), 1024 frames of the 640x480 pyramid per launch, k_blur7 = 1.00 ms:
//   0.64 ms  two ping-pong warpgroups, H through shared memory, issue under `if (thread == 0)`
//   0.79 ms  the same work as a warp-specialised pipeline — slower, which ruled out the dependency chain as the limit
//   ncu + probes: tensor pipe 12 % active, MMA rate ideal when issued alone (64.1 clk per 128x128x32), TMEM reads
//   500-790 B / clk / SM — none of them the limit. The ISSUING THREAD was: inside a divergent branch ptxas cannot prove
//   descriptors / TMEM addresses uniform and wraps every UTCIMMA / UTMALDG in a uniformisation loop (R2UR.BROADCAST +
//   BRA.U.ANY, ~16 instructions, ~170 clk per 64-clk MMA).
//   0.51 ms  the whole warp runs the issue path (warp index and TMEM base broadcast by shfl = provably uniform), only
//            the asynchronous instructions sit under elect.sync, descriptors precomputed
//   0.49 ms  + H in tensor memory (this file)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "orbx_kernels.cuh"

namespace orbx {
namespace {
constexpr int kTX = 128, kTY = 120;
constexpr int kSlab = 4096;
constexpr int kK1 = 5, kK2 = 4;
constexpr int kThreads = 320;
constexpr int kSStages = 3;
constexpr int oT = 0;
constexpr int oW = oT + kK1 * kSlab;
constexpr int oS = oW + kK2 * kSlab;                 // kSStages x (16 KB 128B-swizzled tile + 4 KB 32B-swizzled slab)
constexpr int oOut = oS + kSStages * kK1 * kSlab;    // 2 x [128 rows x 128 B], 128B-swizzled
constexpr int oBar = oOut + 2 * 128 * 128;
constexpr int kSmem = oBar + 256 + 1024;
static_assert(kSmem <= 232448, "shared memory");
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // u8 x u8 -> s32

struct BlurTcParams {
  CUtensorMap src[8];   // box 128 B x 128 rows, SWIZZLE_128B: source columns x0 - 16 .. x0 + 111
  CUtensorMap src2[8];  // box 32 B x 128 rows, SWIZZLE_32B: source columns x0 + 112 .. x0 + 143
  CUtensorMap dst[8];   // box 128 B x 120 rows, SWIZZLE_128B
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
// Bounded spin (~seconds): a protocol error must surface as a CUDA error on the caller's stream, never as a hung device
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  for (uint32_t spins = 0;; spins++) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(sptr(b)), "r"(parity) : "memory");
    if (ok) return;
    if (spins > (1u << 24)) __trap();
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
// 64 accumulator columns, the low 16 bits of two adjacent columns per register
__device__ __forceinline__ void tmem_ld64_pack(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// A from tensor memory (8 cells of 4 K-bytes per lane and K step), B from shared memory
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
// byte offset of element (row, k) of a K-major SWIZZLE_32B operand made of [128 x 32 B] slabs
__device__ __forceinline__ int sw32(int row, int k) {
  return (k >> 5) * kSlab + (row >> 3) * 256 + (row & 7) * 32 + ((((k >> 4) & 1) ^ ((row >> 2) & 1)) << 4) + (k & 15);
}
__device__ __forceinline__ int reflect101_tc(int c, int n) { return c < 0 ? -c : (c >= n ? 2 * n - 2 - c : c); }

}  // namespace

// (outside the unnamed namespace: profilers then show a plain kernel name)
__global__ void __launch_bounds__(kThreads, 1) k_blur_tc(const __grid_constant__ BlurTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sT = smem + oT;
  uint8_t* sW = smem + oW;
  uint8_t* sS = smem + oS;
  uint8_t* sOut = smem + oOut;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);
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
  for (int i = tid; i < (kK1 + kK2) * kSlab / 16; i += kThreads) reinterpret_cast<uint4*>(sT)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    const int taps[7] = {18, 34, 48, 56, 48, 34, 18};
    if (tid < 128) {
      const int x = x0 + tid;
      if (x < w) {
#pragma unroll
        for (int j = 0; j < 7; j++) sT[sw32(tid, reflect101_tc(x + j - 3, w) - (x0 - 16))] += taps[j];
      }
    } else if (tid < 256) {
      const int t = tid - 128, y = y0 + t;
      if (t < kTY && y < h) {
#pragma unroll
        for (int j = 0; j < 7; j++) sW[sw32(t, reflect101_tc(y + j - 3, h) - (y0 - 3))] += taps[j];
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
      if (elect_one()) {
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
      if (elect_one()) {
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
      if (elect_one()) {
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
      tmem_ld64_pack(lane_base, v0);
      tmem_ld64_pack(lane_base + 64, v1);
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
        tmem_ld64_pack(lane_base + 128 + c0, hi);
        tmem_ld64_pack(lane_base + 256 + c0, lo);
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

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tc_encoder() {
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

// The 24 tensor maps depend on the buffers, not on the images: a handle's lanes present the same few (buffers, frames)
// combinations call after call, so the encoded parameter block is kept (24 driver calls ~ 15 us of host time otherwise).
struct MapKey {
  const void *lvl0, *pyr, *blur;
  int64_t fstride0, slab_fstride;
  int64_t off[8];
  int lw[8], lh[8], lpitch[8];
  int pitch0, frames, nlevels, device;
  bool operator==(const MapKey& o) const {
    if (lvl0 != o.lvl0 || pyr != o.pyr || blur != o.blur || fstride0 != o.fstride0 || slab_fstride != o.slab_fstride ||
        pitch0 != o.pitch0 || frames != o.frames || nlevels != o.nlevels || device != o.device)
      return false;
    for (int l = 0; l < nlevels; l++)
      if (off[l] != o.off[l] || lw[l] != o.lw[l] || lh[l] != o.lh[l] || lpitch[l] != o.lpitch[l]) return false;
    return true;
  }
};
struct MapEntry {
  MapKey key;
  BlurTcParams params;
  int classes;
};
std::mutex g_map_mutex;
std::vector<MapEntry> g_map_cache;

bool encode_maps(const Plan& P, const FrameSet& fs, int frames, BlurTcParams* out, int* classes) {
  EncodeTiledFn enc = tc_encoder();
  if (!enc) return false;
  memset(out, 0, sizeof(*out));
  out->nlevels = P.nlevels;
  out->frames = frames;
  *classes = 0;
  for (int l = 0; l < P.nlevels; l++) {
    const LevelPlan& L = P.lv[l];
    if (L.w < 4 || L.h < 4) return false;  // one reflection must bring every tap back inside
    const uint8_t* sbase = l == 0 ? fs.lvl0 : fs.pyr + L.img_off;
    const int64_t spitch = l == 0 ? fs.pitch0 : L.pitch;
    int64_t sfstride = l == 0 ? fs.fstride0 : fs.slab_fstride;
    uint8_t* dbase = fs.blur + L.img_off;
    const int64_t dpitch = L.pitch;
    int64_t dfstride = fs.slab_fstride;
    if (frames == 1) {  // never applied; any legal value
      sfstride = (spitch * L.h + 15) / 16 * 16;
      dfstride = (dpitch * L.h + 15) / 16 * 16;
    }
    if ((reinterpret_cast<uintptr_t>(sbase) & 15) || (spitch & 15) || (sfstride & 15) || spitch < L.w || sfstride <= 0 ||
        (reinterpret_cast<uintptr_t>(dbase) & 15) || (dpitch & 15) || (dfstride & 15) || dfstride <= 0 ||
        dpitch < (L.w + 15) / 16 * 16)
      return false;
    const cuuint64_t dims[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)frames};
    const cuuint64_t sstr[2] = {(cuuint64_t)spitch, (cuuint64_t)sfstride};
    const cuuint64_t dstr[2] = {(cuuint64_t)dpitch, (cuuint64_t)dfstride};
    const cuuint32_t es[3] = {1, 1, 1};
    const cuuint32_t box_s[3] = {128, 128, 1}, box_s2[3] = {32, 128, 1}, box_d[3] = {128, (cuuint32_t)kTY, 1};
    if (enc(&out->src[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(sbase), dims, sstr, box_s, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&out->src2[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(sbase), dims, sstr, box_s2, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&out->dst[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dbase, dims, dstr, box_d, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
    out->w[l] = L.w;
    out->h[l] = L.h;
    *classes += ((L.w + kTX - 1) / kTX) * ((L.h + kTY - 1) / kTY);
  }
  return true;
}
}  // namespace

bool blur_tc_enabled() {
  static const bool on = [] {
    const char* e = getenv("ORBX_BLUR_TC");
    return !(e && e[0] == '0');
  }();
  return on;
}

// false = not applicable here (the caller falls back to k_blur7); true = launched
bool launch_blur_tc(const Plan& P, const FrameSet& fs, int frames, cudaStream_t st) {
  if (!blur_tc_enabled() || P.nlevels > 8 || frames < 1) return false;
  int device = 0;
  cudaGetDevice(&device);
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.lvl0 = fs.lvl0;
  key.pyr = fs.pyr;
  key.blur = fs.blur;
  key.fstride0 = fs.fstride0;
  key.slab_fstride = fs.slab_fstride;
  key.pitch0 = fs.pitch0;
  key.frames = frames;
  key.nlevels = P.nlevels;
  for (int l = 0; l < P.nlevels; l++) {
    key.off[l] = P.lv[l].img_off;
    key.lw[l] = P.lv[l].w;
    key.lh[l] = P.lv[l].h;
    key.lpitch[l] = P.lv[l].pitch;
  }
  key.device = device;
  BlurTcParams params;
  int classes = 0;
  bool found = false;
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    for (const MapEntry& e : g_map_cache)
      if (e.key == key) {
        params = e.params;
        classes = e.classes;
        found = true;
        break;
      }
  }
  if (!found) {
    if (!encode_maps(P, fs, frames, &params, &classes)) return false;
    std::lock_guard<std::mutex> lock(g_map_mutex);
    if (g_map_cache.size() >= 64) g_map_cache.erase(g_map_cache.begin());
    g_map_cache.push_back(MapEntry{key, params, classes});
  }
  // per device, and cheap: set on every launch (a process may drive several GPUs, orbx_extract_batch_multi)
  if (classes < 1 || cudaFuncSetAttribute(k_blur_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
    return false;
  // frames per CTA: enough CTAs for ~4 waves of the 148 SMs, at most 64 (set-up — band matrices, TMEM allocation — costs
  // about two tiles; measured 0.523 / 0.488 / 0.521 ms for 32 / 64 / 128 at 1024 frames)
  int fpc = (int)(((long long)frames * classes + 148 * 4 - 1) / (148 * 4));
  fpc = fpc < 1 ? 1 : (fpc > 64 ? 64 : fpc);
  if (fpc > frames) fpc = frames;
  params.fpc = fpc;
  k_blur_tc<<<dim3(classes, (frames + fpc - 1) / fpc), kThreads, kSmem, st>>>(params);
  return true;
}

}  // namespace orbx
