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

void launch_pyramid(const Plan& P, const FrameSet& fs, const ResizeTab* tab, int frames, cudaStream_t st) {
  for (int l = 1; l < P.nlevels; l++) {
    const LevelPlan& D = P.lv[l];
    dim3 grid((D.pitch / 4 + kResizeThreads - 1) / kResizeThreads, (D.h + kResizeRows - 1) / kResizeRows, frames);
    k_resize<<<grid, kResizeThreads, 0, st>>>(P, fs, tab, l);
  }
}

}  // namespace orbx
