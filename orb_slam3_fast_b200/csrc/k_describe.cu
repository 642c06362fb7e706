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
// Streaming design without shared memory: a thread owns 4 adjacent columns and walks down a strip of kBlurRows output
// rows. Per source row it reads the 3 aligned words x-4 .. x+7, forms the 4 horizontal sums with packed u16x2
// arithmetic (two pixels per IMAD; the sums stay below 65536 so the halves never carry into each other) and pushes them
// into a 7-deep register window; once the window is full every new row yields one 32-bit store of 4 output pixels.
// The row loop is fully unrolled so the window is pure register renaming. Reflect-101 happens in the address (rows)
// or in a byte-wise slow path taken only by the lanes that touch the left / right image edge (columns).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBlurRows = 32;
constexpr int kBlurThreads = 128;

// [b(i), 0, b(i+1), 0] for two adjacent bytes of the 12-byte window (w0 w1 w2); I is compile-time
template <int I>
__device__ __forceinline__ uint32_t bpair(uint32_t w0, uint32_t w1, uint32_t w2) {
  constexpr int word = I >> 2, off = I & 3;
  const uint32_t wa = word == 0 ? w0 : (word == 1 ? w1 : w2);
  if constexpr (off < 3) {
    return __byte_perm(wa, 0u, off | (4 << 4) | ((off + 1) << 8) | (4 << 12));
  } else {
    const uint32_t wb = word == 0 ? w1 : w2;
    const uint32_t t = __byte_perm(wa, wb, 0x0043);  // byte 3 of wa, byte 0 of wb
    return __byte_perm(t, 0u, 0x4140);
  }
}

__global__ void __launch_bounds__(kBlurThreads) k_blur7(const __grid_constant__ Plan P, const FrameSet fs, int l) {
  const LevelPlan& L = P.lv[l];
  const int f = blockIdx.z;
  const int x = 4 * (blockIdx.x * kBlurThreads + threadIdx.x);
  const int y0 = blockIdx.y * kBlurRows;
  const int w = L.w, h = L.h;
  if (x >= w) return;
  int pitch;
  const uint8_t* src = raw_level(P, fs, l, f, &pitch);
  uint8_t* dst = blur_level(P, fs, l, f);
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)pitch) & 3) == 0;
  const bool fast = aligned && x >= 4 && x + 8 <= w;
  int win[7][4];  // horizontal sums of the last 7 source rows
  // the 3 words of row r are requested kBlurAhead iterations before they are used (the loop is fully unrolled, so the
  // queue is register renaming): without this every row pays a full DRAM round trip between load and use
  constexpr int kBlurAhead = 4;
  uint32_t q[kBlurAhead][3];
  auto fetch = [&](int r, uint32_t (&o)[3]) {
    if (y0 + r - 6 >= h) return;  // this source row completes no output row
    const int ys = reflect101(y0 - 3 + r, h);
    const uint8_t* row = src + (int64_t)ys * pitch;
    if (fast) {
      const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row + x - 4);
      o[0] = __ldg(r32);
      o[1] = __ldg(r32 + 1);
      o[2] = __ldg(r32 + 2);
    } else {
      // bytes x-4 .. x+7 with reflect-101 columns; columns beyond w+2 only feed outputs beyond w
      uint32_t b[12];
#pragma unroll
      for (int k = 0; k < 12; k++) {
        const int xx = x - 4 + k;
        b[k] = (xx >= -3 && xx < w + 3) ? row[reflect101(xx, w)] : 0;
      }
      o[0] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
      o[1] = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
      o[2] = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
    }
  };
#pragma unroll
  for (int r = 0; r < kBlurAhead; r++) fetch(r, q[r]);
#pragma unroll
  for (int r = 0; r < kBlurRows + 6; r++) {
    const int yo = y0 + r - 6;  // output row completed by this source row
    if (yo >= h) break;
    const uint32_t w0 = q[r % kBlurAhead][0], w1 = q[r % kBlurAhead][1], w2 = q[r % kBlurAhead][2];
    if (r + kBlurAhead < kBlurRows + 6) fetch(r + kBlurAhead, q[r % kBlurAhead]);
    // window byte i = column x - 4 + i; output pixel k reads bytes k+1 .. k+7
    const uint32_t o0 = bpair<1>(w0, w1, w2), e1 = bpair<2>(w0, w1, w2), o1 = bpair<3>(w0, w1, w2);
    const uint32_t e2 = bpair<4>(w0, w1, w2), o2 = bpair<5>(w0, w1, w2), e3 = bpair<6>(w0, w1, w2);
    const uint32_t o3 = bpair<7>(w0, w1, w2), e4 = bpair<8>(w0, w1, w2), o4 = bpair<9>(w0, w1, w2);
    const uint32_t h01 = 18u * (o0 + o3) + 34u * (e1 + e3) + 48u * (o1 + o2) + 56u * e2;  // pixels 0, 1
    const uint32_t h23 = 18u * (o1 + o4) + 34u * (e2 + e4) + 48u * (o2 + o3) + 56u * e3;  // pixels 2, 3
#pragma unroll
    for (int j = 0; j < 6; j++)
#pragma unroll
      for (int k = 0; k < 4; k++) win[j][k] = win[j + 1][k];
    win[6][0] = (int)(h01 & 0xffff);
    win[6][1] = (int)(h01 >> 16);
    win[6][2] = (int)(h23 & 0xffff);
    win[6][3] = (int)(h23 >> 16);
    if (r >= 6) {
      uint32_t packed = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t acc = 32768u + 18u * (uint32_t)(win[0][k] + win[6][k]) + 34u * (uint32_t)(win[1][k] + win[5][k]) +
                             48u * (uint32_t)(win[2][k] + win[4][k]) + 56u * (uint32_t)win[3][k];
        packed |= (acc >> 16) << (8 * k);
      }
      *reinterpret_cast<uint32_t*>(dst + (int64_t)yo * L.pitch + x) = packed;
    }
  }
}

void launch_blur(const Plan& P, const FrameSet& fs, int frames, cudaStream_t st) {
  for (int l = 0; l < P.nlevels; l++) {
    const LevelPlan& L = P.lv[l];
    dim3 grid(((L.w + 3) / 4 + kBlurThreads - 1) / kBlurThreads, (L.h + kBlurRows - 1) / kBlurRows, frames);
    k_blur7<<<grid, kBlurThreads, 0, st>>>(P, fs, l);
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
