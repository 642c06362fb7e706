// k_describe.cu — everything after the quadtree in ORBextractor::operator() (src/ORBextractor.cc:1059-1105):
//   k_blur7     cv::GaussianBlur(level, 7x7, sigma 2, REFLECT_101) (:1074-1076), OpenCV's fixed-point path
//   k_assemble  output row of every keypoint: levels ascending, "mono" rows from the front, rows whose scaled x lies in
//               the lapping area from the back (:1083-1101)
//   k_describe  IC_Angle (:75-99, :471-488) on the raw level + computeOrbDescriptor (:102-147) on the blurred level,
//               one warp per keypoint, writing the final cv::KeyPoint record and descriptor row
#include "orbx_kernels.cuh"
#include "orbx_quadtree.h"

namespace orbx {

// ---------------------------------------------------------------------------------------------------------------
// 7x7 Gaussian, Q0.8 taps {18,34,48,56,48,34,18} on both axes, 16-bit row sums, one rounding: (sum + 32768) >> 16.
// Block = 32x8 threads, output tile 128 x 16; the 22 needed rows of horizontal sums are kept in shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBlurTW = 128, kBlurTH = 16;

__device__ __forceinline__ int hsum7(int a, int b, int c, int d, int e, int f, int g) {
  return 18 * (a + g) + 34 * (b + f) + 48 * (c + e) + 56 * d;
}

__global__ void __launch_bounds__(256) k_blur7(const __grid_constant__ Plan P, const FrameSet fs, int l) {
  __shared__ uint16_t H[kBlurTH + 6][kBlurTW];
  const LevelPlan& L = P.lv[l];
  const int f = blockIdx.z;
  const int x0 = blockIdx.x * kBlurTW, y0 = blockIdx.y * kBlurTH;
  const int tx = threadIdx.x, ty = threadIdx.y;
  int pitch;
  const uint8_t* src = raw_level(P, fs, l, f, &pitch);
  const int w = L.w, h = L.h;
  const int x = x0 + 4 * tx;
  for (int rr = ty; rr < kBlurTH + 6; rr += 8) {
    const int y = reflect101(y0 - 3 + rr, h);
    const uint8_t* row = src + (int64_t)y * pitch;
    uint8_t b[10];
    if (x >= 4 && x + 7 < w && ((reinterpret_cast<uintptr_t>(row) & 3) == 0)) {
      const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row + x - 4);
      const uint32_t w0 = r32[0], w1 = r32[1], w2 = r32[2];
      b[0] = (w0 >> 8) & 0xff; b[1] = (w0 >> 16) & 0xff; b[2] = w0 >> 24;
      b[3] = w1 & 0xff; b[4] = (w1 >> 8) & 0xff; b[5] = (w1 >> 16) & 0xff; b[6] = w1 >> 24;
      b[7] = w2 & 0xff; b[8] = (w2 >> 8) & 0xff; b[9] = (w2 >> 16) & 0xff;
    } else {
#pragma unroll
      for (int k = 0; k < 10; k++) {
        const int xx = x - 3 + k;
        b[k] = xx < w + 3 ? row[reflect101(xx, w)] : 0;  // columns past w+2 only feed outputs past w
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
      H[rr][4 * tx + k] = (uint16_t)hsum7(b[k], b[k + 1], b[k + 2], b[k + 3], b[k + 4], b[k + 5], b[k + 6]);
  }
  __syncthreads();
  uint8_t* dst = blur_level(P, fs, l, f);
  for (int r = ty; r < kBlurTH; r += 8) {
    const int y = y0 + r;
    if (y >= h || x >= L.pitch) continue;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = 4 * tx + k;
      const uint32_t acc = 32768u + 18u * (H[r][c] + H[r + 6][c]) + 34u * (H[r + 1][c] + H[r + 5][c]) +
                           48u * (H[r + 2][c] + H[r + 4][c]) + 56u * H[r + 3][c];
      packed |= (acc >> 16) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(dst + (int64_t)y * L.pitch + x) = packed;
  }
}

void launch_blur(const Plan& P, const FrameSet& fs, int frames, cudaStream_t st) {
  for (int l = 0; l < P.nlevels; l++) {
    const LevelPlan& L = P.lv[l];
    dim3 grid((L.w + kBlurTW - 1) / kBlurTW, (L.h + kBlurTH - 1) / kBlurTH, frames);
    k_blur7<<<grid, dim3(32, 8), 0, st>>>(P, fs, l);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Output rows. One warp per frame; ballot prefix sums keep the encounter order the serial reference has.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_assemble(const __grid_constant__ Plan P, const WorkSet ws, const OutSet out,
                                                 int lap0, int lap1) {
  const int f = blockIdx.x, lane = threadIdx.x;
  int total = 0;
  for (int l = 0; l < P.nlevels; l++) total += ws.lvl_n[f * P.nlevels + l];
  const bool fits = total <= out.cap;
  int mono = 0, stereo = total - 1;
  const float flap0 = (float)lap0, flap1 = (float)lap1;
  for (int l = 0; l < P.nlevels; l++) {
    const LevelPlan& L = P.lv[l];
    const int n = ws.lvl_n[f * P.nlevels + l];
    const uint32_t* kp = ws.lvl_kp + (int64_t)f * P.kps_per_frame + L.kp_base;
    int32_t* dst = ws.dst + (int64_t)f * P.kps_per_frame + L.kp_base;
    for (int base = 0; base < n; base += 32) {
      const int p = base + lane;
      bool is_st = false, valid = p < n;
      if (valid) {
        float x = (float)(cand_x(kp[p]) + kMinBorder);
        if (l != 0) x = fmul(x, L.scale);           // keypoint->pt *= scale          :1086
        is_st = x >= flap0 && x <= flap1;           // inclusive lapping test          :1088-1089
      }
      const unsigned m_st = __ballot_sync(0xffffffffu, valid && is_st);
      const unsigned m_mo = __ballot_sync(0xffffffffu, valid && !is_st);
      const unsigned lt = (1u << lane) - 1u;
      if (valid) {
        const int d = is_st ? stereo - __popc(m_st & lt) : mono + __popc(m_mo & lt);
        dst[p] = fits ? d : -1;
      }
      stereo -= __popc(m_st);
      mono += __popc(m_mo);
    }
  }
  if (lane == 0) {
    out.n[f] = total;
    out.mono[f] = mono;  // the reference's return value (monoIndex, :1105)
    out.status[f] = fits ? 0 : -2;
  }
}

void launch_assemble(const Plan& P, const WorkSet& ws, const OutSet& out, int lap0, int lap1, int frames,
                     cudaStream_t st) {
  k_assemble<<<frames, 32, 0, st>>>(P, ws, out, lap0, lap1);
}

// ---------------------------------------------------------------------------------------------------------------
// Orientation + descriptor, one warp per keypoint.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDescWarps = 4;

__global__ void __launch_bounds__(kDescWarps * 32)
k_describe(const __grid_constant__ Plan P, const FrameSet fs, const WorkSet ws, const OutSet out,
           const int8_t* __restrict__ pattern) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * kDescWarps + warp;
  const int f = blockIdx.y;
  if (e >= P.kps_per_frame) return;
  int l = 0;
  while (l + 1 < P.nlevels && P.lv[l + 1].kp_base <= e) l++;
  const LevelPlan& L = P.lv[l];
  const int p = e - L.kp_base;
  if (p >= ws.lvl_n[f * P.nlevels + l]) return;
  const int dst = ws.dst[(int64_t)f * P.kps_per_frame + e];
  if (dst < 0) return;
  const uint32_t cw = ws.lvl_kp[(int64_t)f * P.kps_per_frame + e];
  const int X = cand_x(cw) + kMinBorder, Y = cand_y(cw) + kMinBorder;  // :867-868

  // ---- IC_Angle: m10 = sum u*I, m01 = sum v*I over the radius-15 disc; lane = column u + 15 ----
  int pitch;
  const uint8_t* img = raw_level(P, fs, l, f, &pitch);
  const uint8_t* center = img + (int64_t)Y * pitch + X;
  const int u = lane - kHalfPatch;
  const int au = u < 0 ? -u : u;
  int m10 = 0, m01 = 0;
  if (lane < 31) {
#pragma unroll
    for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
      const int d = P.umax[v < 0 ? -v : v];
      if (au <= d) {
        const int val = center[v * pitch + u];
        m10 += u * val;
        m01 += v * val;
      }
    }
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
  const int4* pat4 = reinterpret_cast<const int4*>(pattern) + lane * 2;
  const int4 q0 = pat4[0], q1 = pat4[1];
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
