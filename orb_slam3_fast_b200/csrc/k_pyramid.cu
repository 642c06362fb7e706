// k_pyramid.cu — ORBextractor::ComputePyramid (src/ORBextractor.cc:1108-1145): level l = cv::resize(level l-1,
// INTER_LINEAR) (:1122). OpenCV's 8-bit bilinear is fixed point (11-bit coefficients, two-step rounding); the tables
// are built on the host by orbx::axis_table. The 19-px reflect-101 frame the reference adds around every level
// (:1129-1143) is never read by the hot path, so the device pyramid is border-less; the border is
// synthesised only by orbx_download_pyramid for the host mirror of mvImagePyramid.
//
// One launch per level (7 dependent steps) over all frames of the batch. Thread = 4 consecutive destination pixels,
// one 32-bit store; the two source rows are gathered through L1 (a 32-lane warp touches ~154 consecutive source bytes).
#include "orbx_kernels.cuh"

namespace orbx {

__global__ void __launch_bounds__(128) k_resize(const __grid_constant__ Plan P, const FrameSet fs,
                                                const ResizeTab* __restrict__ tab, int l) {
  const LevelPlan& D = P.lv[l];
  const LevelPlan& S = P.lv[l - 1];
  const int d0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const int f = blockIdx.z;
  if (d0 >= D.pitch) return;
  int spitch;
  const uint8_t* src = raw_level(P, fs, l - 1, f, &spitch);
  uint8_t* dst = fs.pyr + (int64_t)f * fs.slab_fstride + D.img_off + (int64_t)y * D.pitch;
  const ResizeTab ty = tab[D.ytab_off + y];
  int sy0 = ty.ofs, sy1 = ty.ofs + 1;  // rows are clipped, the coefficients kept (resizeGeneric_Invoker)
  sy0 = sy0 < 0 ? 0 : (sy0 >= S.h ? S.h - 1 : sy0);
  sy1 = sy1 < 0 ? 0 : (sy1 >= S.h ? S.h - 1 : sy1);
  const uint8_t* r0 = src + (int64_t)sy0 * spitch;
  const uint8_t* r1 = src + (int64_t)sy1 * spitch;
  const int b0 = ty.c0, b1 = ty.c1;
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int d = d0 + k;
    if (d < D.w) {
      const ResizeTab tx = tab[D.xtab_off + d];
      const int s = tx.ofs;
      const int s1 = s + 1 < S.w ? s + 1 : S.w - 1;
      const int h0 = r0[s] * tx.c0 + r0[s1] * tx.c1;
      const int h1 = r1[s] * tx.c0 + r1[s1] * tx.c1;
      int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      v = v < 0 ? 0 : (v > 255 ? 255 : v);
      packed |= (uint32_t)v << (8 * k);
    }
  }
  *reinterpret_cast<uint32_t*>(dst + d0) = packed;  // pitch is a multiple of 64: the padding is written as zeros
}

void launch_pyramid(const Plan& P, const FrameSet& fs, const ResizeTab* tab, int frames, cudaStream_t st) {
  for (int l = 1; l < P.nlevels; l++) {
    const LevelPlan& D = P.lv[l];
    dim3 block(128);
    dim3 grid((D.pitch / 4 + 127) / 128, D.h, frames);
    k_resize<<<grid, block, 0, st>>>(P, fs, tab, l);
  }
}

}  // namespace orbx
