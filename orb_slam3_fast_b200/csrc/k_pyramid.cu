// k_pyramid.cu — ORBextractor::ComputePyramid (src/ORBextractor.cc:1108-1145): level l = cv::resize(level l-1,
// INTER_LINEAR) (:1122). OpenCV's 8-bit bilinear is fixed point (11-bit coefficients, two-step rounding); the tables
// are built on the host by orbx::axis_table. The 19-px reflect-101 frame the reference adds around every level
// (:1129-1143) is never read by the hot path, so the device pyramid is border-less; the border is
// synthesised only by orbx_download_pyramid for the host mirror of mvImagePyramid.
//
// One launch per level (7 dependent steps) over all frames of the batch. Streaming design, no shared memory:
// a thread owns 4 consecutive destination columns and a strip of kResizeRows destination rows, which it produces one
// after the other. Its column taps (offset + the coefficient pair packed for IDP.2A) are loaded once; per source row
// it reads the 3 aligned words that cover its 4 x 2 source bytes (requested one destination row ahead), PRMT-selects
// each byte pair and forms c0*b0 + c1*b1 with one dp2a. The second source row's horizontal results stay in registers:
// the next destination row usually starts on it. One 32-bit store per 4 pixels. (The first version walked the SOURCE
// rows and emitted destination rows from a while loop: 42 instructions per pixel, issue bound at 1.4 TB/s.)
#include <cuda.h>  // CUtensorMap types; the encoder comes from cudaGetDriverEntryPoint

#include "orbx_kernels.cuh"

namespace orbx {

constexpr int kResizeRows = 8;
constexpr int kResizeThreads = 128;

__global__ void __launch_bounds__(kResizeThreads)
k_resize(const __grid_constant__ Plan P, const FrameSet fs, const ResizeTab* __restrict__ tab, int l) {
  const LevelPlan& D = P.lv[l];
  const LevelPlan& S = P.lv[l - 1];
  const int d0 = (blockIdx.x * kResizeThreads + threadIdx.x) * 4;
  const int y_begin = blockIdx.y * kResizeRows;
  const int f = blockIdx.z;
  if (d0 >= D.pitch) return;
  int spitch;
  const uint8_t* src = raw_level(P, fs, l - 1, f, &spitch);
  uint8_t* dst = fs.pyr + (int64_t)f * fs.slab_fstride + D.img_off;
  const int dpitch = D.pitch, sh1 = S.h - 1, sw1 = S.w - 1;

  // ---- column taps of the 4 destination pixels ----
  int s[4];
  uint32_t coef[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int d = d0 + k;
    if (d < D.w) {
      const ResizeTab tx = tab[D.xtab_off + d];
      s[k] = tx.ofs;
      coef[k] = (uint32_t)(uint16_t)tx.c0 | ((uint32_t)(uint16_t)tx.c1 << 16);
    } else {
      s[k] = d0 < D.w ? s[0] : 0;
      coef[k] = 0;  // padding columns are written as zeros
    }
  }
  const int base = s[0] & ~3;
  // fast path: the 12 bytes [base, base + 12) hold every tap and lie inside the row; rows are word aligned. The
  // window is first shifted right by the thread's misalignment s[0] & 3 (two funnel shifts per row), after which all
  // 8 tap bytes of the 4 pixels sit in 8 consecutive bytes and one PRMT per pixel picks its pair.
  // The last thread of a row may find its third word past the row end (level 0 is read with the caller's pitch, which
  // can equal the width): that word then only holds right-hand taps of weight 0, so it is simply not loaded. (Before
  // this, that one lane took the byte-load path and its whole warp — one in five — waited for it: 52 % of the stalls.)
  bool fast = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)spitch) & 3) == 0 && base + 8 <= spitch;
  const bool ld2 = base + 12 <= spitch;
  const uint32_t mis = (uint32_t)(s[0] - base) * 8u;
  uint32_t sel[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int o = s[k] - s[0];  // byte offset of the left tap in the shifted window; the right tap is o + 1
    if (o < 0 || o + 1 > 7) fast = false;
    sel[k] = (uint32_t)(o & 7) | ((uint32_t)((o + 1) & 7) << 4) | 0x4400u;  // bytes 2,3 of the result: don't care
  }
  const uint8_t* srcb = src + base;

  struct Row3 {
    uint32_t w0, w1, w2;
  };
  auto fetch = [&](int sy) {
    Row3 r{0u, 0u, 0u};
    if (fast) {
      const uint32_t* r32 = reinterpret_cast<const uint32_t*>(srcb + (int64_t)sy * spitch);
      r.w0 = __ldg(r32);
      r.w1 = __ldg(r32 + 1);
      if (ld2) r.w2 = __ldg(r32 + 2);
    }
    return r;
  };
  // horizontal pass of one source row -> 4 ints, already >> 4 (the vertical pass only uses them that way)
  auto hrow = [&](int sy, const Row3& r, int (&h)[4]) {
    if (fast) {
      const uint32_t a0 = __funnelshift_r(r.w0, r.w1, mis), a1 = __funnelshift_r(r.w1, r.w2, mis);
#pragma unroll
      for (int k = 0; k < 4; k++)
        h[k] = (int)__dp2a_lo(coef[k], __byte_perm(a0, a1, sel[k]), 0u) >> 4;  // (c0 * b0 + c1 * b1) >> 4
    } else {
      const uint8_t* row = src + (int64_t)sy * spitch;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int s1 = s[k] + 1 < sw1 ? s[k] + 1 : sw1;
        h[k] = ((int)row[s[k]] * (int)(coef[k] & 0xffff) + (int)row[s1] * (int)(coef[k] >> 16)) >> 4;
      }
    }
  };
  auto clip = [&](int v) { return v < 0 ? 0 : (v > sh1 ? sh1 : v); };  // rows are clipped, the coefficients kept

  const int y_end = min(y_begin + kResizeRows, D.h);
  const ResizeTab* ytab = tab + D.ytab_off;
  // Destination rows one after the other. The two source rows of row y are (a0, a1); at a scale >= 1 row y + 1
  // starts on a1 four times out of five, in which case its first horizontal pass is the previous row's second one
  // (kept in registers) and only one new source row is requested. All row decisions are block uniform (every thread
  // of the block works on the same rows), so none of this diverges. The source words of row y + 1 are requested
  // before row y is computed.
  ResizeTab ty = ytab[y_begin];
  int a0 = clip(ty.ofs), a1 = clip(ty.ofs + 1);
  Row3 q0 = fetch(a0), q1 = fetch(a1);
  int hp[4] = {0, 0, 0, 0};
  int prev = -1;
  uint8_t* dptr = dst + (int64_t)y_begin * dpitch + d0;
  for (int y = y_begin; y < y_end; y++) {
    ResizeTab tn = ty;
    int n0 = a0, n1 = a1;
    Row3 p0 = q0, p1 = q1;
    if (y + 1 < y_end) {
      tn = ytab[y + 1];
      n0 = clip(tn.ofs);
      n1 = clip(tn.ofs + 1);
      if (n0 != a1) p0 = fetch(n0);
      if (n1 != n0) p1 = fetch(n1);
    }
    int h0[4], h1[4];
    if (a0 == prev) {
#pragma unroll
      for (int k = 0; k < 4; k++) h0[k] = hp[k];
    } else {
      hrow(a0, q0, h0);
    }
    if (a1 == a0) {  // both taps on one row (clipped at the image top / bottom): the coefficients are kept
#pragma unroll
      for (int k = 0; k < 4; k++) h1[k] = h0[k];
    } else {
      hrow(a1, q1, h1);
    }
    const int b0 = ty.c0, b1 = ty.c1;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
      v[k] = (uint32_t)((((b0 * h0[k]) >> 16) + ((b1 * h1[k]) >> 16) + 2) >> 2);  // in [0, 255]: taps sum to 2048
    *reinterpret_cast<uint32_t*>(dptr) = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
    dptr += dpitch;
#pragma unroll
    for (int k = 0; k < 4; k++) hp[k] = h1[k];
    prev = a1;
    ty = tn;
    a0 = n0;
    a1 = n1;
    q0 = p0;
    q1 = p1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA variant (the one that normally runs). The LDG kernel above is latency bound in the step (1.4 TB/s with ~16 KB of
// loads in flight per SM, and removing a quarter of its instructions did not move its in-step time): here a block
// first pulls the SOURCE rectangle its 128 x 32 destination tile needs into shared memory with ONE
// cp.async.bulk.tensor box (box_w x box_h bytes, starting at the 16-byte boundary below the first tap — an unaligned
// innermost coordinate faults, tools/ubench/tma_probe.cu), so ~8 KB per resident block are in flight without costing
// registers; the arithmetic is the same as above, reading 32-bit words from the tile. Scale factors whose rectangle
// is wider than the 256-byte box limit (> ~1.75) take the LDG kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRtColThreads = 32;   // x 4 destination columns = 128 per block: the source rectangle fits ONE box
constexpr int kRtRowGroups = 4;     // x kResizeRows destination rows = 32 per block
constexpr int kRtCols = kRtColThreads * 4, kRtRows = kRtRowGroups * kResizeRows;
constexpr int kRtHead = 128;        // mbarrier in front of the tile
constexpr int kRtMaxBoxes = 1;

struct ResizeMaps {
  CUtensorMap lv[8];  // u8 [frames][h][w] view of the SOURCE level l - 1 at index l - 1; box = (box_w, box_h, 1)
};

__global__ void __launch_bounds__(kRtColThreads * kRtRowGroups)
k_resize_tma(const __grid_constant__ Plan P, const __grid_constant__ ResizeMaps maps, const FrameSet fs,
             const ResizeTab* __restrict__ tab, int l, int box_w, int box_h) {
  extern __shared__ __align__(128) uint8_t smem[];
  const LevelPlan& D = P.lv[l];
  const LevelPlan& S = P.lv[l - 1];
  const int tx = threadIdx.x % kRtColThreads, tr = threadIdx.x / kRtColThreads;
  const int dblock = blockIdx.x * kRtCols;  // < D.w: pitch is w rounded up to 64, the tile width is a multiple of it
  const int d0 = dblock + tx * 4;
  const int yb = blockIdx.y * kRtRows;
  const int f = blockIdx.z;
  const ResizeTab* xtab = tab + D.xtab_off;
  const ResizeTab* ytab = tab + D.ytab_off;
  const int sh1 = S.h - 1, sw1 = S.w - 1;
  auto clip = [&](int v) { return v < 0 ? 0 : (v > sh1 ? sh1 : v); };  // rows are clipped, the coefficients kept
  const int xa = xtab[dblock].ofs & ~15;
  const int ys0 = clip(ytab[yb].ofs);
  uint8_t* tile = smem + kRtHead - xa;  // tile[c + row * box_w] = source column c of tile row `row`
  {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_w * box_h) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"((uint32_t)__cvta_generic_to_shared(smem + kRtHead)), "l"(reinterpret_cast<uint64_t>(&maps.lv[l - 1])),
          "r"(xa), "r"(ys0), "r"(f), "r"(bar)
          : "memory");
    }
    __syncthreads();
    uint32_t ok;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(bar) : "memory");
    } while (!ok);
  }
  const int y_begin = yb + tr * kResizeRows;
  if (d0 >= D.pitch || y_begin >= D.h) return;
  uint8_t* dst = fs.pyr + (int64_t)f * fs.slab_fstride + D.img_off;
  const int dpitch = D.pitch;

  // ---- column taps of the 4 destination pixels (as in k_resize) ----
  int s[4];
  uint32_t coef[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int d = d0 + k;
    if (d < D.w) {
      const ResizeTab t = xtab[d];
      s[k] = t.ofs;
      coef[k] = (uint32_t)(uint16_t)t.c0 | ((uint32_t)(uint16_t)t.c1 << 16);
    } else {
      s[k] = d0 < D.w ? s[0] : xa;
      coef[k] = 0;  // padding columns are written as zeros
    }
  }
  const int base = s[0] & ~3;
  bool fast = true;
  const uint32_t mis = (uint32_t)(s[0] - base) * 8u;
  uint32_t sel[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int o = s[k] - s[0];
    if (o < 0 || o + 1 > 7) fast = false;
    sel[k] = (uint32_t)(o & 7) | ((uint32_t)((o + 1) & 7) << 4) | 0x4400u;
  }
  const uint8_t* pw = tile + base;
  auto hrow = [&](int sy, int (&h)[4]) {
    const int ro = (sy - ys0) * box_w;
    if (fast) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(pw + ro);
      const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
      const uint32_t a0 = __funnelshift_r(w0, w1, mis), a1 = __funnelshift_r(w1, w2, mis);
#pragma unroll
      for (int k = 0; k < 4; k++)
        h[k] = (int)__dp2a_lo(coef[k], __byte_perm(a0, a1, sel[k]), 0u) >> 4;  // (c0 * b0 + c1 * b1) >> 4
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++)
        h[k] = ((int)tile[ro + s[k]] * (int)(coef[k] & 0xffff) +
                (int)tile[ro + (s[k] + 1 < sw1 ? s[k] + 1 : sw1)] * (int)(coef[k] >> 16)) >> 4;
    }
  };
  const int y_end = min(y_begin + kResizeRows, D.h);
  int hp[4] = {0, 0, 0, 0};
  int prev = -1;
  uint8_t* dptr = dst + (int64_t)y_begin * dpitch + d0;
  for (int y = y_begin; y < y_end; y++) {
    const ResizeTab ty = ytab[y];
    const int a0 = clip(ty.ofs), a1 = clip(ty.ofs + 1);
    int h0[4], h1[4];
    if (a0 == prev) {
#pragma unroll
      for (int k = 0; k < 4; k++) h0[k] = hp[k];
    } else {
      hrow(a0, h0);
    }
    if (a1 == a0) {
#pragma unroll
      for (int k = 0; k < 4; k++) h1[k] = h0[k];
    } else {
      hrow(a1, h1);
    }
    const int b0 = ty.c0, b1 = ty.c1;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
      v[k] = (uint32_t)((((b0 * h0[k]) >> 16) + ((b1 * h1[k]) >> 16) + 2) >> 2);
    *reinterpret_cast<uint32_t*>(dptr) = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
    dptr += dpitch;
#pragma unroll
    for (int k = 0; k < 4; k++) hp[k] = h1[k];
    prev = a1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::cvtColor(..., COLOR_{BGR,RGB,BGRA,RGBA}2GRAY) on 8-bit images (src/Tracking.cc:1394-1412): OpenCV's fixed point
// gray = (B * 3735 + G * 19235 + R * 9798 + 16384) >> 15. SURVEY.md §8(f) rank 4 (the image front-end). Pure
// streaming: a thread turns 4 pixels (three or four aligned words) into one output word; the only kernel of the
// library that is bound by HBM rather than by instruction issue.
// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// cv::remap(src, dst, mapx, mapy, INTER_LINEAR), CV_8UC1 / CV_32FC1 maps / BORDER_CONSTANT 0: the stereo rectification
// of System::TrackStereo (src/System.cc:293-294). OpenCV's fixed point: coordinates in 1/32 px, exact 5-bit weights,
// (sum + 512) >> 10. A thread produces 4 neighbouring destination pixels (one float4 of each map, 16 gathered bytes,
// one output word); the maps are shared by every frame of the batch, so they stay in L2.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRemapFrames = 4;  // frames per thread: the map values, tap addresses and weights are computed once for them

__global__ void __launch_bounds__(256)
k_remap_linear(const uint8_t* __restrict__ src, int sw, int sh, int sstride, int64_t sfstride,
               const float* __restrict__ mapx, const float* __restrict__ mapy, int dw, int dh, uint8_t* __restrict__ dst,
               int dstride, int64_t dfstride, int vec, int frames) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y, f0 = blockIdx.z * kRemapFrames;
  if (x >= dw) return;
  const int64_t mo = (int64_t)y * dw + x;
  float mx[4], my[4];
  const int n = min(4, dw - x);
  const bool wide = vec && n == 4;
  if (wide) {
    const float4 a = *reinterpret_cast<const float4*>(mapx + mo), b = *reinterpret_cast<const float4*>(mapy + mo);
    mx[0] = a.x; mx[1] = a.y; mx[2] = a.z; mx[3] = a.w;
    my[0] = b.x; my[1] = b.y; my[2] = b.z; my[3] = b.w;
  } else {
    for (int k = 0; k < 4; k++) {
      mx[k] = k < n ? mapx[mo + k] : 0.f;
      my[k] = k < n ? mapy[mo + k] : 0.f;
    }
  }
  int off[4], fxs[4], fys[4];
  unsigned ok[4];  // bit 0..3: taps (0,0) (0,1) (1,0) (1,1) inside the source
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int sx = __float2int_rn(fmul(mx[k], 32.f)), sy = __float2int_rn(fmul(my[k], 32.f));  // cvRound(map * 32)
    const int ix = sx >> 5, iy = sy >> 5;
    fxs[k] = sx & 31;
    fys[k] = sy & 31;
    const bool x0 = ix >= 0 && ix < sw, x1 = ix + 1 >= 0 && ix + 1 < sw, y0 = iy >= 0 && iy < sh, y1 = iy + 1 >= 0 && iy + 1 < sh;
    ok[k] = (unsigned)(x0 && y0) | ((unsigned)(x1 && y0) << 1) | ((unsigned)(x0 && y1) << 2) | ((unsigned)(x1 && y1) << 3);
    // clamp the address of a fully outside tap so that `off` stays a valid int; it is never dereferenced then
    off[k] = ok[k] ? iy * sstride + ix : 0;
  }
  for (int df = 0; df < kRemapFrames && f0 + df < frames; df++) {
    const uint8_t* s = src + (f0 + df) * sfstride;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint8_t* r0 = s + off[k];
      const uint8_t* r1 = r0 + sstride;
      const int p00 = (ok[k] & 1u) ? r0[0] : 0, p01 = (ok[k] & 2u) ? r0[1] : 0;
      const int p10 = (ok[k] & 4u) ? r1[0] : 0, p11 = (ok[k] & 8u) ? r1[1] : 0;
      const int top = (32 - fxs[k]) * p00 + fxs[k] * p01, bot = (32 - fxs[k]) * p10 + fxs[k] * p11;
      out |= (uint32_t)(((32 - fys[k]) * top + fys[k] * bot + 512) >> 10) << (8 * k);
    }
    uint8_t* d = dst + (f0 + df) * dfstride + (int64_t)y * dstride + x;
    if (wide) {
      *reinterpret_cast<uint32_t*>(d) = out;
    } else {
      for (int k = 0; k < n; k++) d[k] = (uint8_t)(out >> (8 * k));
    }
  }
}

int launch_remap_linear(const uint8_t* src, int sw, int sh, int sstride, int64_t sfstride, const float* mapx,
                        const float* mapy, int dw, int dh, uint8_t* dst, int dstride, int64_t dfstride, int frames,
                        cudaStream_t st) {
  const int vec = (dw % 4 == 0) && ((reinterpret_cast<uintptr_t>(mapx) | reinterpret_cast<uintptr_t>(mapy)) & 15) == 0 &&
                  ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)dstride | (uintptr_t)dfstride) & 3) == 0;
  dim3 grid(((dw + 3) / 4 + 255) / 256, dh, (frames + kRemapFrames - 1) / kRemapFrames);
  k_remap_linear<<<grid, 256, 0, st>>>(src, sw, sh, sstride, sfstride, mapx, mapy, dw, dh, dst, dstride, dfstride, vec,
                                       frames);
  return 0;
}

constexpr int kCvtRows = 4;  // rows per thread: their loads are all issued before the first pixel is computed

template <int kChannels>
__global__ void __launch_bounds__(256)
k_cvt_gray(const uint8_t* __restrict__ src, int w, int h, int sstride, int64_t sfstride, uint8_t* __restrict__ dst,
           int dstride, int64_t dfstride, int rgb, int aligned) {
  const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
  const int y0 = blockIdx.y * (4 * kCvtRows) + (threadIdx.x >> 6);  // rows y0, y0 + 4, y0 + 8, y0 + 12
  const int f = blockIdx.z;
  if (x >= w) return;
  const uint8_t* s = src + f * sfstride + (int64_t)x * kChannels;
  uint8_t* d = dst + f * dfstride + x;
  const uint32_t c0 = rgb ? 9798u : 3735u, c2 = rgb ? 3735u : 9798u;  // weight of the first / third channel
  if (aligned && x + 4 <= w) {
    uint32_t raw[kCvtRows][4];
#pragma unroll
    for (int k = 0; k < kCvtRows; k++) {
      const int y = y0 + 4 * k;
      if (y < h) {
        const uint8_t* p = s + (int64_t)y * sstride;
        if (kChannels == 4) {
          const uint4 q = *reinterpret_cast<const uint4*>(p);
          raw[k][0] = q.x; raw[k][1] = q.y; raw[k][2] = q.z; raw[k][3] = q.w;
        } else {
          raw[k][0] = reinterpret_cast<const uint32_t*>(p)[0];
          raw[k][1] = reinterpret_cast<const uint32_t*>(p)[1];
          raw[k][2] = reinterpret_cast<const uint32_t*>(p)[2];
          raw[k][3] = 0;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kCvtRows; k++) {
      const int y = y0 + 4 * k;
      if (y >= h) continue;
      uint32_t px[4];  // per pixel: byte 0 = first channel, byte 1 = green, byte 2 = third channel
      if (kChannels == 4) {
        px[0] = raw[k][0]; px[1] = raw[k][1]; px[2] = raw[k][2]; px[3] = raw[k][3];
      } else {
        px[0] = raw[k][0];
        px[1] = __funnelshift_r(raw[k][0], raw[k][1], 24);
        px[2] = __funnelshift_r(raw[k][1], raw[k][2], 16);
        px[3] = raw[k][2] >> 8;
      }
      uint32_t out = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t v = (px[q] & 0xff) * c0 + ((px[q] >> 8) & 0xff) * 19235u + ((px[q] >> 16) & 0xff) * c2 + 16384u;
        out |= (v >> 15) << (8 * q);
      }
      *reinterpret_cast<uint32_t*>(d + (int64_t)y * dstride) = out;
    }
  } else {
    for (int k = 0; k < kCvtRows; k++) {
      const int y = y0 + 4 * k;
      if (y >= h) continue;
      for (int q = 0; q < 4 && x + q < w; q++) {
        const uint8_t* p = s + (int64_t)y * sstride + q * kChannels;
        d[(int64_t)y * dstride + q] = (uint8_t)(((uint32_t)p[0] * c0 + (uint32_t)p[1] * 19235u + (uint32_t)p[2] * c2 + 16384u) >> 15);
      }
    }
  }
}

int launch_cvt_gray(const uint8_t* src, int w, int h, int sstride, int64_t sfstride, int channels, int rgb, uint8_t* dst,
                    int dstride, int64_t dfstride, int frames, cudaStream_t st) {
  if (channels != 3 && channels != 4) return -1;
  const int aligned = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)sstride | (uintptr_t)sfstride) & (channels == 4 ? 15 : 3)) == 0 &&
                      ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)dstride | (uintptr_t)dfstride) & 3) == 0;
  dim3 grid(((w + 3) / 4 + 63) / 64, (h + 4 * kCvtRows - 1) / (4 * kCvtRows), frames);
  if (channels == 3) k_cvt_gray<3><<<grid, 256, 0, st>>>(src, w, h, sstride, sfstride, dst, dstride, dfstride, rgb, aligned);
  else k_cvt_gray<4><<<grid, 256, 0, st>>>(src, w, h, sstride, sfstride, dst, dstride, dfstride, rgb, aligned);
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn resize_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct ResizeTile {
  int box_w, box_h, nbox;
};
// Source rectangle of a kRtCols x kRtRows destination tile, from the tap formula s(d) = floor((d + 0.5) * scale - 0.5).
static ResizeTile resize_tile(const LevelPlan& S, const LevelPlan& D) {
  const double sx = (double)S.w / D.w, sy = (double)S.h / D.h;
  // columns: up to 15 in front (aligned start), the taps of the tile's pixels, the right tap, the 12-byte word window
  const int need_w = 15 + (int)ceil((kRtCols - 1) * sx) + 2 + 12 + 1;
  const int need_h = (int)ceil((kRtRows - 1) * sy) + 3;
  ResizeTile t;
  t.nbox = (need_w + 255) / 256;
  t.box_w = round_up(need_w, 16);
  t.box_h = need_h;
  return t;
}

void launch_pyramid(const Plan& P, const FrameSet& fs, const ResizeTab* tab, int frames, cudaStream_t st) {
  // tensor maps of the source levels 0 .. nlevels - 2; any level that cannot be described sends the whole chain to
  // the LDG kernel (only a caller-owned level 0 with an odd base / pitch / frame stride can do that)
  ResizeMaps M;
  ResizeTile T[kMaxLevels];
  bool tma = resize_encode_tiled() != nullptr && P.nlevels - 1 <= 8;
  for (int l = 1; l < P.nlevels && tma; l++) {
    const LevelPlan& S = P.lv[l - 1];
    T[l] = resize_tile(S, P.lv[l]);
    const uint8_t* base = l == 1 ? fs.lvl0 : fs.pyr + S.img_off;
    const int64_t pitch = l == 1 ? fs.pitch0 : S.pitch;
    int64_t fstride = l == 1 ? fs.fstride0 : fs.slab_fstride;
    if (frames == 1) fstride = (pitch * S.h + 15) / 16 * 16;  // never applied
    if (T[l].nbox > kRtMaxBoxes || T[l].box_h > 256 || kRtHead + T[l].box_w * T[l].box_h > 48 * 1024 || (reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) ||
        (fstride & 15) || pitch <= 0 || fstride <= 0) {
      tma = false;
      break;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)S.w, (cuuint64_t)S.h, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
    const cuuint32_t box[3] = {(cuuint32_t)T[l].box_w, (cuuint32_t)T[l].box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (resize_encode_tiled()(&M.lv[l - 1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      tma = false;
  }
  for (int l = 1; l < P.nlevels; l++) {
    const LevelPlan& D = P.lv[l];
    if (tma) {
      const size_t smem = kRtHead + (size_t)T[l].box_w * T[l].box_h;
      dim3 grid((D.pitch + kRtCols - 1) / kRtCols, (D.h + kRtRows - 1) / kRtRows, frames);
      k_resize_tma<<<grid, kRtColThreads * kRtRowGroups, smem, st>>>(P, M, fs, tab, l, T[l].box_w, T[l].box_h);
    } else {
      dim3 grid((D.pitch / 4 + kResizeThreads - 1) / kResizeThreads, (D.h + kResizeRows - 1) / kResizeRows, frames);
      k_resize<<<grid, kResizeThreads, 0, st>>>(P, fs, tab, l);
    }
  }
}

}  // namespace orbx
