// k_pyramid.cu — ORBextractor::ComputePyramid (src/ORBextractor.cc:1108-1145): level l = cv::resize(level l-1,
// INTER_LINEAR) (:1122). OpenCV's 8-bit bilinear is fixed point (11-bit coefficients, two-step rounding); the tables
// are built on the host by orbx::axis_table. The 19-px reflect-101 frame the reference adds around every level
// (:1129-1143) is never read by the hot path, so the device pyramid is border-less; the border is
// synthesised only by orbx_download_pyramid for the host mirror of mvImagePyramid.
//
// One launch per level (7 dependent steps) over all frames of the batch. Streaming design, no shared memory:
// a thread owns 4 consecutive destination columns and walks kResizeRows destination rows. Its column taps
// (offset + the coefficient pair packed for IDP.2A) are loaded once; per source row it reads the 3 aligned words
// that cover its 4 x 2 source bytes, PRMT-selects each byte pair and forms c0*b0 + c1*b1 with one dp2a. Horizontal
// results of a source row are kept for the next destination row (consecutive rows share a source row ~80% of the
// time at scale 1.2). One 32-bit store per 4 pixels.
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
  // fast path: the 12 bytes [base, base + 12) hold every tap and lie inside the row; rows are word aligned
  bool fast = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)spitch) & 3) == 0 && base + 12 <= spitch;
  uint32_t sel[4];
  bool hi[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int o = s[k] - base;  // byte offset of the left tap; the right tap is o + 1 (its weight is 0 when clamped)
    if (o < 0 || o + 1 > 11) fast = false;
    hi[k] = o > 6;
    const int oo = hi[k] ? o - 4 : o;
    sel[k] = (uint32_t)(oo & 7) | ((uint32_t)((oo + 1) & 7) << 4) | 0x4400u;  // bytes 2,3 of the result: don't care
  }

  // horizontal pass of one source row -> 4 ints
  auto hrow = [&](int sy, int (&h)[4]) {
    const uint8_t* row = src + (int64_t)sy * spitch;
    if (fast) {
      const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row + base);
      const uint32_t w0 = r32[0], w1 = r32[1], w2 = r32[2];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t pair = __byte_perm(hi[k] ? w1 : w0, hi[k] ? w2 : w1, sel[k]);
        h[k] = (int)__dp2a_lo(coef[k], pair, 0u);  // c0 * b0 + c1 * b1
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int s1 = s[k] + 1 < S.w ? s[k] + 1 : S.w - 1;
        h[k] = (int)row[s[k]] * (int)(coef[k] & 0xffff) + (int)row[s1] * (int)(coef[k] >> 16);
      }
    }
  };

  // the two most recent source rows (indices are warp-uniform: they depend on y only)
  int ia = -1, ib = -1;
  int ha[4] = {0, 0, 0, 0}, hb[4] = {0, 0, 0, 0};
  for (int r = 0; r < kResizeRows; r++) {
    const int y = y_begin + r;
    if (y >= D.h) break;
    const ResizeTab ty = tab[D.ytab_off + y];
    int sy0 = ty.ofs, sy1 = ty.ofs + 1;  // rows are clipped, the coefficients kept (resizeGeneric_Invoker)
    sy0 = sy0 < 0 ? 0 : (sy0 >= S.h ? S.h - 1 : sy0);
    sy1 = sy1 < 0 ? 0 : (sy1 >= S.h ? S.h - 1 : sy1);
    const int b0 = ty.c0, b1 = ty.c1;
    int n0[4], n1[4];
    if (sy0 == ia) {
#pragma unroll
      for (int k = 0; k < 4; k++) n0[k] = ha[k];
    } else if (sy0 == ib) {
#pragma unroll
      for (int k = 0; k < 4; k++) n0[k] = hb[k];
    } else {
      hrow(sy0, n0);
    }
    if (sy1 == sy0) {
#pragma unroll
      for (int k = 0; k < 4; k++) n1[k] = n0[k];
    } else if (sy1 == ib) {
#pragma unroll
      for (int k = 0; k < 4; k++) n1[k] = hb[k];
    } else {
      hrow(sy1, n1);
    }
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int v = (((b0 * (n0[k] >> 4)) >> 16) + ((b1 * (n1[k] >> 4)) >> 16) + 2) >> 2;
      v = v < 0 ? 0 : (v > 255 ? 255 : v);
      packed |= (uint32_t)v << (8 * k);
      ha[k] = n0[k];
      hb[k] = n1[k];
    }
    ia = sy0;
    ib = sy1;
    *reinterpret_cast<uint32_t*>(dst + (int64_t)y * D.pitch + d0) = packed;
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
