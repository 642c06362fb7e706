// k_describe.cu — everything after the quadtree in ORBextractor::operator() (src/ORBextractor.cc:1059-1105):
//   k_blur7     cv::GaussianBlur(level, 7x7, sigma 2, REFLECT_101) (:1074-1076), OpenCV's fixed-point path
//   k_describe  IC_Angle (:75-99, :471-488) on the raw level + computeOrbDescriptor (:102-147) on the blurred level,
//               one warp per keypoint, writing the final cv::KeyPoint record and descriptor row at its output row:
//               levels ascending, "mono" rows from the front, rows whose scaled x lies in the lapping area from the
//               back (:1083-1101) — the per-level ranks come from k_quadtree, the level offsets are summed here
#include <cuda.h>  // CUtensorMap types; the encoder comes from cudaGetDriverEntryPoint

#include "orbx_kernels.cuh"
#include "orbx_quadtree.h"

namespace orbx {

// ---------------------------------------------------------------------------------------------------------------
// 7x7 Gaussian, Q0.8 taps {18,34,48,56,48,34,18} on both axes, 16-bit row sums, one rounding: (sum + 32768) >> 16.
//
// ONE launch covers every level of every frame. The unit of work is a warp task = (level, 128-column strip,
// kBlurRows-row strip); a lane owns 4 adjacent columns and walks down the strip. Streaming, no shared memory:
//  * per source row a lane reads the 3 aligned words that hold columns x-4 .. x+7. The 4 horizontal sums are 8 DP4A
//    (4 taps each, the 8th coefficient is 0) on byte groups cut out of the window with 6 funnel shifts;
//  * the sums go into a 7-deep register window (the row loop is fully unrolled, so the window is register renaming);
//    once it is full every new source row yields one 32-bit store of 4 output pixels;
//  * the words of row r are requested kBlurAhead iterations before they are used;
//  * reflect-101 costs nothing in the common case: rows are reflected in the address (warp uniform); the one or two
//    lanes of a row that touch the left / right image edge patch their window with three PRMTs whose selectors are
//    computed once per task — no byte loads, no divergent slow path (the first version of this kernel spent 2/3 of
//    its time in edge warps).
// Where the words come from is the template parameter: kBlurTma (the normal case) — the warp's whole source rectangle
// (160 x 38 bytes: the strip, a 3-row halo above and below, the aligned words left and right) is brought into shared
// memory by ONE cp.async.bulk.tensor box per warp before the row loop, so the loop's loads are LDS and the bytes in
// flight per SM (~6 KB per resident warp) no longer depend on registers (the LDG variant sat at 1.6 TB/s in the step,
// stalled on its prefetch queue); kBlurLdg — the same loop on aligned global words; kBlurBytes — level 0 read in
// place from a caller buffer whose base or pitch is not word aligned (byte loads).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBlurRows = 32;
constexpr int kBlurWarps = 4;
constexpr int kBlurAhead = 4;
constexpr int kBlurBytes = 0, kBlurLdg = 1, kBlurTma = 2;
constexpr int kBlurBoxW = 160, kBlurBoxH = kBlurRows + 6;             // bytes x rows per warp
constexpr int kBlurTile = (kBlurBoxW * kBlurBoxH + 127) / 128 * 128;  // TMA destinations are 128-byte aligned
constexpr int kBlurHead = 128;                                        // one mbarrier per warp

struct BlurMaps {
  CUtensorMap lv[8];  // u8 [frames][h][w] view of every raw level; box = (160, 38, 1)
};

__device__ __forceinline__ int blur_strips_x(int w) { return (w + 127) >> 7; }
__device__ __forceinline__ int blur_strips_y(int h) { return (h + kBlurRows - 1) / kBlurRows; }

template <int kMode>
__global__ void __launch_bounds__(kBlurWarps * 32)
k_blur7(const __grid_constant__ Plan P, const __grid_constant__ BlurMaps maps, const FrameSet fs) {
  constexpr bool kAligned = kMode != kBlurBytes;
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  int t = blockIdx.x * kBlurWarps + (threadIdx.x >> 5);
  const int f = blockIdx.y;
  int l = 0;
  for (;; l++) {
    if (l == P.nlevels) return;
    const int n = blur_strips_x(P.lv[l].w) * blur_strips_y(P.lv[l].h);
    if (t < n) break;
    t -= n;
  }
  const LevelPlan& L = P.lv[l];
  const int w = L.w, h = L.h;
  const int sx = blur_strips_x(w);
  const int ty = t / sx, tx = t - ty * sx;
  const int x = 4 * (tx * 32 + lane);
  const int y0 = ty * kBlurRows;
  const uint32_t* tile32 = nullptr;
  if constexpr (kMode == kBlurTma) {
    // box origin: 16 bytes left of the strip (the lane's first word is x - 4; the start must be 16-byte aligned) and 3
    // rows above it; everything outside the level — including negative coordinates — arrives as zeros
    const int warp = threadIdx.x >> 5;
    uint8_t* tile = smem + kBlurHead + warp * kBlurTile;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem + 8 * warp);
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBlurBoxW * kBlurBoxH) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"((uint32_t)__cvta_generic_to_shared(tile)), "l"(reinterpret_cast<uint64_t>(&maps.lv[l])),
          "r"(128 * tx - 16), "r"(y0 - 3), "r"(f), "r"(bar)
          : "memory");
    }
    __syncwarp();
    uint32_t ok;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(bar) : "memory");
    } while (!ok);
    tile32 = reinterpret_cast<const uint32_t*>(tile) + lane + 3;  // word of column x - 4 in tile row 0
  }
  if (x >= w) return;
  int pitch;
  const uint8_t* src = raw_level(P, fs, l, f, &pitch);
  uint8_t* dst = blur_level(P, fs, l, f);

  // ---- edge patch: window byte i holds column x - 4 + i; columns outside [0, w) take their reflect-101 source, which
  //      always lies inside the two neighbouring words of the same window ----
  const bool ld0 = x >= 4, ld2 = x + 8 <= pitch;  // words that exist (word 1 always does)
  const bool edge = x < 4 || x + 8 > w;
  uint32_t sel[3] = {0x3210u, 0x7654u, 0x7654u};
  bool hi[3] = {false, false, true};  // operands of the PRMT: (w0, w1) or (w1, w2)
  if (edge) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      int j[4], lo = 12, top = -1;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int c = x - 4 + 4 * k + b;
        j[b] = -1;                                   // don't care
        if (c >= -3 && c <= w + 2) {
          j[b] = reflect101(c, w) - (x - 4);
          lo = min(lo, j[b]);
          top = max(top, j[b]);
        }
      }
      const bool up = top >= 8;                      // then every source is >= 4 (checked for every w in tests)
      const int base = up ? 4 : 0;
      uint32_t s_ = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) s_ |= (uint32_t)((j[b] < 0 ? (lo < 12 ? lo : 4 * k + b) : j[b]) - base) << (4 * b);
      sel[k] = top < 0 ? (k == 0 ? 0x3210u : 0x7654u) : s_;
      hi[k] = top < 0 ? k == 2 : up;
    }
  }

  int win[7][4];
  uint32_t q[kBlurAhead][3];
  // source rows past the bottom reflection range only feed output rows >= h, which are never stored: clamp them so
  // that the row loop needs no early exit (the unrolled loop stays branch free)
  const uint8_t* srcx = src + x;
  const int h2 = 2 * h - 2;
  auto fetch = [&](int r, uint32_t (&o)[3]) {
    const int v = abs(min(y0 - 3 + r, h + 2));
    const int ys = min(v, h2 - v);  // reflect-101 of a row in [-3, h + 2], branch free
    if constexpr (kMode == kBlurTma) {
      const uint32_t* r32 = tile32 + (ys - (y0 - 3)) * (kBlurBoxW / 4);
      o[0] = r32[0];
      o[1] = r32[1];
      o[2] = r32[2];
      return;
    }
    const uint8_t* row = src + (int64_t)ys * pitch;
    if (kAligned) {
      const uint32_t* r32 = reinterpret_cast<const uint32_t*>(srcx + (int64_t)ys * pitch);
      o[0] = ld0 ? __ldg(r32 - 1) : 0u;
      o[1] = __ldg(r32);
      o[2] = ld2 ? __ldg(r32 + 1) : 0u;
    } else {
      uint32_t b[12];
#pragma unroll
      for (int k = 0; k < 12; k++) {
        const int xx = x - 4 + k;
        b[k] = (xx >= 0 && xx < w) ? row[xx] : 0;
      }
      o[0] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
      o[1] = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
      o[2] = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
    }
  };
  constexpr uint32_t kTapsA = 18u | (34u << 8) | (48u << 16) | (56u << 24);
  constexpr uint32_t kTapsB = 48u | (34u << 8) | (18u << 16);
  const int dpitch = L.pitch;
  uint8_t* dptr = dst + (int64_t)y0 * dpitch + x;  // output row y0 + r - 6 is completed by source row r
  int rows_left = h - y0;
#pragma unroll
  for (int r = 0; r < kBlurAhead; r++) fetch(r, q[r]);
#pragma unroll
  for (int r = 0; r < kBlurRows + 6; r++) {
    uint32_t w0 = q[r % kBlurAhead][0], w1 = q[r % kBlurAhead][1], w2 = q[r % kBlurAhead][2];
    if (r + kBlurAhead < kBlurRows + 6) fetch(r + kBlurAhead, q[r % kBlurAhead]);
    if (edge) {
      const uint32_t n0 = __byte_perm(hi[0] ? w1 : w0, hi[0] ? w2 : w1, sel[0]);
      const uint32_t n1 = __byte_perm(hi[1] ? w1 : w0, hi[1] ? w2 : w1, sel[1]);
      const uint32_t n2 = __byte_perm(hi[2] ? w1 : w0, hi[2] ? w2 : w1, sel[2]);
      w0 = n0;
      w1 = n1;
      w2 = n2;
    }
#pragma unroll
    for (int j = 0; j < 6; j++)
#pragma unroll
      for (int k = 0; k < 4; k++) win[j][k] = win[j + 1][k];
    // output pixel k reads window bytes k+1 .. k+7: group A = bytes k+1..k+4, group B = bytes k+5..k+8 (tap 8 = 0).
    // Every row sum starts at 128: the vertical taps add up to 256, so the 7 rows carry the +32768 of the final rounding
    win[6][0] = (int)__dp4a(__funnelshift_r(w0, w1, 8), kTapsA, __dp4a(__funnelshift_r(w1, w2, 8), kTapsB, 128u));
    win[6][1] = (int)__dp4a(__funnelshift_r(w0, w1, 16), kTapsA, __dp4a(__funnelshift_r(w1, w2, 16), kTapsB, 128u));
    win[6][2] = (int)__dp4a(__funnelshift_r(w0, w1, 24), kTapsA, __dp4a(__funnelshift_r(w1, w2, 24), kTapsB, 128u));
    win[6][3] = (int)__dp4a(w1, kTapsA, __dp4a(w2, kTapsB, 128u));
    if (r >= 6) {
      uint32_t acc[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
        acc[k] = 18u * (uint32_t)(win[0][k] + win[6][k]) + 34u * (uint32_t)(win[1][k] + win[5][k]) +
                 48u * (uint32_t)(win[2][k] + win[4][k]) + 56u * (uint32_t)win[3][k];
      // byte 2 of every accumulator = (sum + 32768) >> 16
      const uint32_t packed = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
      if (r - 6 < rows_left) *reinterpret_cast<uint32_t*>(dptr) = packed;
      dptr += dpitch;
    }
  }
}

typedef CUresult (*BlurEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_blur_maps(const Plan& P, const FrameSet& fs, int frames, BlurMaps* M) {
  static BlurEncodeFn enc = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<BlurEncodeFn>(p);
  }();
  if (!enc || P.nlevels > 8) return false;
  for (int l = 0; l < P.nlevels; l++) {
    const uint8_t* base = l == 0 ? fs.lvl0 : fs.pyr + P.lv[l].img_off;
    const int64_t pitch = l == 0 ? fs.pitch0 : P.lv[l].pitch;
    int64_t fstride = l == 0 ? fs.fstride0 : fs.slab_fstride;
    if (frames == 1) fstride = (pitch * P.lv[l].h + 15) / 16 * 16;  // never applied
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (fstride & 15) || pitch <= 0 || fstride <= 0)
      return false;
    const cuuint64_t dims[3] = {(cuuint64_t)P.lv[l].w, (cuuint64_t)P.lv[l].h, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
    const cuuint32_t box[3] = {(cuuint32_t)kBlurBoxW, (cuuint32_t)kBlurBoxH, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (enc(&M->lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

void launch_blur(const Plan& P, const FrameSet& fs, int frames, cudaStream_t st) {
  if (launch_blur_tc(P, fs, frames, st)) return;  // the tensor-core form (k_blur_tc.cu); below: the CUDA-core fallbacks
  int tasks = 0;
  for (int l = 0; l < P.nlevels; l++) tasks += ((P.lv[l].w + 127) / 128) * ((P.lv[l].h + kBlurRows - 1) / kBlurRows);
  dim3 grid((tasks + kBlurWarps - 1) / kBlurWarps, frames);
  BlurMaps M;
  if (make_blur_maps(P, fs, frames, &M)) {
    static const bool carve = [] {
      cudaFuncSetAttribute(k_blur7<kBlurTma>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      return true;
    }();
    (void)carve;
    k_blur7<kBlurTma><<<grid, kBlurWarps * 32, kBlurHead + kBlurWarps * kBlurTile, st>>>(P, M, fs);
    return;
  }
  memset(&M, 0, sizeof(M));
  // levels >= 1 are owned buffers (256-byte aligned, pitch multiple of 64); level 0 is the caller's
  const bool aligned = ((reinterpret_cast<uintptr_t>(fs.lvl0) | (uintptr_t)fs.pitch0 | (uintptr_t)fs.fstride0) & 3) == 0;
  if (aligned) k_blur7<kBlurLdg><<<grid, kBlurWarps * 32, 0, st>>>(P, M, fs);
  else k_blur7<kBlurBytes><<<grid, kBlurWarps * 32, 0, st>>>(P, M, fs);
}

// ---------------------------------------------------------------------------------------------------------------
// Orientation + descriptor, one warp per keypoint.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDescWarps = 4;

// Measured dead end (round 2): splitting this kernel into moments (warp per keypoint) / angles (THREAD per keypoint: the
// fastAtan2 + FP64 sincosf chain that every lane repeats here) / descriptors (warp per keypoint) cut the instruction count
// by a third and made the stage SLOWER (2.37 -> 2.75 ms per 2048 frames): the chain runs in the shadow of the other
// warps' gathers, while two more passes over the keypoint records do not.
__global__ void __launch_bounds__(kDescWarps * 32)
k_describe(const __grid_constant__ Plan P, const FrameSet fs, const WorkSet ws, const OutSet out,
           const int8_t* __restrict__ pattern) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * kDescWarps + warp;
  const int f = blockIdx.y;
  if (e >= P.kps_per_frame) return;
  // level of entry e: the last level whose first entry is <= e (lane k looks at level k)
  const int l = __popc(__ballot_sync(0xffffffffu, lane < P.nlevels && P.lv[lane < P.nlevels ? lane : 0].kp_base <= e)) - 1;
  const LevelPlan& L = P.lv[l];
  const int p = e - L.kp_base;
  // ---- output row (:1083-1101): monoIndex counts up from 0, stereoIndex down from N - 1, levels ascending.
  //      Lane k holds level k's counts; the sums over all levels / the levels below l are warp reductions ----
  // the entry's own record does not depend on the counts: requested first, so that its latency overlaps theirs
  const int rk = ws.dst[(int64_t)f * P.kps_per_frame + e];
  const uint32_t cw = ws.lvl_kp[(int64_t)f * P.kps_per_frame + e];
  int n_k = 0, s_k = 0;
  if (lane < P.nlevels) {
    n_k = ws.lvl_n[f * P.nlevels + lane];
    s_k = ws.lvl_st[f * P.nlevels + lane];
  }
  const int total = __reduce_add_sync(0xffffffffu, n_k);
  const int st_total = __reduce_add_sync(0xffffffffu, s_k);
  const int mono_before = __reduce_add_sync(0xffffffffu, lane < l ? n_k - s_k : 0);
  const int st_before = __reduce_add_sync(0xffffffffu, lane < l ? s_k : 0);
  const int n_here = __shfl_sync(0xffffffffu, n_k, l);
  const bool fits = total <= out.cap;
  if (e == 0 && lane == 0) {
    out.n[f] = total;
    out.mono[f] = total - st_total;  // the reference's return value (monoIndex, :1105)
    out.status[f] = fits ? 0 : -2;
  }
  if (p >= n_here || !fits) return;
  const int dst = (rk & 0x40000000) ? total - 1 - (st_before + (rk & 0x3fffffff)) : mono_before + rk;
  const int X = cand_x(cw) + kMinBorder, Y = cand_y(cw) + kMinBorder;  // :867-868

  // ---- IC_Angle: m10 = sum u*I, m01 = sum v*I over the radius-15 disc; lane = column u + 15 ----
  int pitch;
  const uint8_t* img = raw_level(P, fs, l, f, &pitch);
  const uint8_t* center = img + (int64_t)Y * pitch + X;
  // lane = column u + 15; a row v takes part where |u| <= umax[|v|]. Since u is fixed per lane, m10 = u * (sum of the
  // lane's pixels); the row pointer is bumped by the pitch (no per-row multiply) and umax is a compile-time table.
  const int u = lane - kHalfPatch;
  const int au = u < 0 ? -u : u;
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    constexpr int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};  // :456-468, HALF_PATCH 15
    const uint8_t* rowp = center - (int64_t)kHalfPatch * pitch + u;
    int colsum = 0;
#pragma unroll
    for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
      if (au <= kUmax[v < 0 ? -v : v]) {
        const int val = *rowp;
        colsum += val;
        m01 += v * val;
      }
      rowp += pitch;
    }
    m10 = u * colsum;
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // ---- rBRIEF: lane i produces byte i (tests 8i .. 8i+7) ----
  const float factorPI = 0x1.1df46ap-6f;  // (float)(CV_PI / 180.f)                      :100
  float a, b;
  sincosf_glibc(fmul(angle, factorPI), &a, &b);  // a = cos, b = sin                    :106-107
  const uint8_t* bimg = blur_level(P, fs, l, f);
  const int bp = L.pitch;
  const uint8_t* bc = bimg + (int64_t)Y * bp + X;
  // 32 bytes of pattern per lane: 8 tests x (x0, y0, x1, y1) as int8. (A float table was tried: 4x the L1 traffic per
  // warp made the kernel 40 % slower than converting here.)
  const int4* pat4 = reinterpret_cast<const int4*>(pattern) + lane * 2;
  const int4 q0 = __ldg(pat4), q1 = __ldg(pat4 + 1);
  const int words[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
  uint32_t byte = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int wv = words[j];
    const float x0 = (float)(int8_t)(wv & 0xff), y0 = (float)(int8_t)((wv >> 8) & 0xff);
    const float x1 = (float)(int8_t)((wv >> 16) & 0xff), y1 = (float)(int8_t)((wv >> 24) & 0xff);
    // GET_VALUE(idx) = center[cvRound(x*b + y*a) * step + cvRound(x*a - y*b)]          :112-114
    const int r0 = cv_round(fadd(fmul(x0, b), fmul(y0, a))), c0 = cv_round(fsub(fmul(x0, a), fmul(y0, b)));
    const int r1 = cv_round(fadd(fmul(x1, b), fmul(y1, a))), c1 = cv_round(fsub(fmul(x1, a), fmul(y1, b)));
    const int t0 = bc[r0 * bp + c0], t1 = bc[r1 * bp + c1];
    byte |= (uint32_t)(t0 < t1) << j;
  }
  out.desc[((int64_t)f * out.cap + dst) * ORBX_DESC_BYTES + lane] = (uint8_t)byte;
  if (lane == 0) {
    orbx_kp k;
    k.x = (float)X;
    k.y = (float)Y;
    if (l != 0) {
      k.x = fmul(k.x, L.scale);
      k.y = fmul(k.y, L.scale);
    }
    k.size = (float)L.patch;
    k.angle = angle;
    k.response = (float)cand_s(cw);
    k.octave = l;
    k.class_id = -1;
    out.kps[(int64_t)f * out.cap + dst] = k;
  }
}

void launch_describe(const Plan& P, const FrameSet& fs, const WorkSet& ws, const OutSet& out, const int8_t* pattern,
                     int frames, cudaStream_t st) {
  dim3 grid((P.kps_per_frame + kDescWarps - 1) / kDescWarps, frames);
  k_describe<<<grid, kDescWarps * 32, 0, st>>>(P, fs, ws, out, pattern);
}

}  // namespace orbx
