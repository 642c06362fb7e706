// orbref.cpp — CPU ORACLE (test infrastructure; see orbref.h for the pinning status).
//
// Serial-order restatement of the reference front-end. Every function cites the reference file:line it follows
// (paths relative to the reference checkout, hellovuong/ORB_SLAM3_FAST @ 6255e16). OpenCV / glibc / libstdc++
// arithmetic that the reference merely *calls* is restated from the published algorithms and pinned against
// cv2 4.13.0 by tests/test_oracle_primitives.py.
//
// Float discipline: the reference is compiled without -march=native (CMakeLists.txt:13-18) => plain SSE2 float
// math, no FMA contraction. This file must be compiled with -ffp-contract=off and without -ffast-math.
#include "orbref.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

// The hot loop of FAST uses SSE2 intrinsics (baseline x86-64: the reference is built without -march=native,
// CMakeLists.txt:13-18, and an OpenCV binary runs its SSE2 / dispatched kernels either way) with OpenCV's own early exit
// per 16-pixel block; integer arithmetic, same bytes as the plain loop it replaces (kept as the non-x86 fallback).
// Measured dead end: GCC function multi-versioning (target_clones "avx2") of the three primitives made the extractor 37 %
// SLOWER on the build host (per-cell calls through the resolver, AVX <-> SSE transitions), so there is no AVX2 variant.
#if defined(__SSE2__)
#include <emmintrin.h>
#define ORBREF_PRIMITIVES_KIND "the oracle's cv2-pinned restatements: SSE2 FAST with OpenCV's per-16-pixel early exit, compiler-vectorised (SSE2) resize / blur"
#else
#define ORBREF_PRIMITIVES_KIND "the oracle's cv2-pinned restatements (portable build, no SIMD intrinsics)"
#endif
#define ORBREF_SIMD_CLONES

namespace {

constexpr int kPatch = 31;      // PATCH_SIZE        src/ORBextractor.cc:71
constexpr int kHalfPatch = 15;  // HALF_PATCH_SIZE   :72
constexpr int kEdge = 19;       // EDGE_THRESHOLD    :73

// cvRound(float/double): SSE cvtss2si / cvtsd2si under the default rounding mode = round half to even.
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }

const int8_t kPattern[256 * 4] = {
#include "../orb_slam3_fast_b200/csrc/orb_pattern.inc"
};

inline int reflect101(int p, int len) {
  // cv::borderInterpolate(p, len, BORDER_REFLECT_101)
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

// ---------------------------------------------------------------------------------------------------------------
// cv::resize(u8, INTER_LINEAR) — OpenCV imgproc resize.cpp: fixed-point bilinear, INTER_RESIZE_COEF_BITS = 11.
// Called at src/ORBextractor.cc:1122.
// ---------------------------------------------------------------------------------------------------------------
struct AxisTab {
  std::vector<int> ofs;
  std::vector<short> c0, c1;
};

AxisTab axis_table(int ssize, int dsize, bool clamp_like_x) {
  AxisTab t;
  t.ofs.resize(dsize);
  t.c0.resize(dsize);
  t.c1.resize(dsize);
  const double inv_scale = (double)dsize / ssize;
  const double scale = 1. / inv_scale;
  for (int d = 0; d < dsize; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (clamp_like_x) {
      if (s < 0) { f = 0; s = 0; }
      if (s >= ssize - 1) { f = 0; s = ssize - 1; }
    }
    t.ofs[d] = s;
    t.c0[d] = (short)std::min(std::max(cv_round((1.f - f) * 2048.f), -32768), 32767);
    t.c1[d] = (short)std::min(std::max(cv_round(f * 2048.f), -32768), 32767);
  }
  return t;
}

ORBREF_SIMD_CLONES void resize_linear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
  AxisTab tx = axis_table(sw, dw, true), ty = axis_table(sh, dh, false);
  std::vector<int> r0(dw), r1(dw);
  int have0 = -1000000, have1 = -1000000;  // source rows held by r0 / r1 (OpenCV keeps its horizontal rows the same way)
  auto hrow = [&](int sy, std::vector<int>& out) {
    sy = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);  // rows are clipped, beta is kept (resizeGeneric_Invoker)
    const uint8_t* S = src + (size_t)sy * sstride;
    for (int d = 0; d < dw; d++) {
      int s = tx.ofs[d];
      int s1 = std::min(s + 1, sw - 1);
      out[d] = S[s] * tx.c0[d] + S[s1] * tx.c1[d];
    }
  };
  for (int y = 0; y < dh; y++) {
    const int sy = ty.ofs[y];
    if (sy == have1) {  // the lower row of the previous output row is this one's upper row
      r0.swap(r1);
      have0 = have1;
      have1 = -1000000;
    }
    if (sy != have0) {
      hrow(sy, r0);
      have0 = sy;
    }
    if (sy + 1 != have1) {
      hrow(sy + 1, r1);
      have1 = sy + 1;
    }
    const int b0 = ty.c0[y], b1 = ty.c1[y];
    uint8_t* D = dst + (size_t)y * dstride;
    for (int x = 0; x < dw; x++) {
      int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
      D[x] = (uint8_t)std::min(std::max(v, 0), 255);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::GaussianBlur(u8, Size(7,7), 2, 2, BORDER_REFLECT_101) — OpenCV smooth.dispatch.cpp fixed-point path:
// Q0.8 kernel {18,34,48,56,48,34,18}, 16-bit horizontal sums, one rounding at the end. src/ORBextractor.cc:1075-1076.
// ---------------------------------------------------------------------------------------------------------------
ORBREF_SIMD_CLONES void gauss7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  // same arithmetic as before (Q0.8 taps, 16-bit row sums, one rounding), arranged so that gcc vectorises both passes:
  // a reflect-padded copy of each row, then 7 shifted adds; the vertical pass walks 7 row pointers.
  std::vector<uint16_t> H((size_t)w * h);
  std::vector<uint8_t> pad((size_t)w + 6);
  for (int y = 0; y < h; y++) {
    const uint8_t* S = src + (size_t)y * sstride;
    for (int i = 0; i < 3; i++) {
      pad[i] = S[reflect101(i - 3, w)];
      pad[(size_t)w + 3 + i] = S[reflect101(w + i, w)];
    }
    memcpy(pad.data() + 3, S, (size_t)w);
    const uint8_t* P = pad.data();
    uint16_t* Hr = H.data() + (size_t)y * w;
    for (int x = 0; x < w; x++)
      Hr[x] = (uint16_t)(18 * (P[x] + P[x + 6]) + 34 * (P[x + 1] + P[x + 5]) + 48 * (P[x + 2] + P[x + 4]) + 56 * P[x + 3]);
  }
  for (int y = 0; y < h; y++) {
    uint8_t* D = dst + (size_t)y * dstride;
    const uint16_t* r[7];
    for (int j = 0; j < 7; j++) r[j] = H.data() + (size_t)reflect101(y + j - 3, h) * w;
    for (int x = 0; x < w; x++) {
      const uint32_t acc = 32768u + 18u * ((uint32_t)r[0][x] + r[6][x]) + 34u * ((uint32_t)r[1][x] + r[5][x]) +
                           48u * ((uint32_t)r[2][x] + r[4][x]) + 56u * (uint32_t)r[3][x];
      D[x] = (uint8_t)(acc >> 16);
    }
  }
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101 [+ISOLATED]) — src/ORBextractor.cc:1129-1143
void border101(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int border) {
  for (int y = -border; y < h + border; y++) {
    const uint8_t* S = src + (size_t)reflect101(y, h) * sstride;
    uint8_t* D = dst + (size_t)(y + border) * dstride;
    for (int x = -border; x < w + border; x++) D[x + border] = S[reflect101(x, w)];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cv::FAST(TYPE_9_16, nonmaxSuppression = true) — OpenCV features2d fast.cpp / fast_score.cpp.
// For centre v and ring p_k: d_k = v - p_k. m = max over the 16 arcs of 9 consecutive ring pixels of
// max(min d_k, min -d_k). Corner at threshold T <=> m > T; cornerScore = m - 1 (the largest threshold at which the
// pixel is still a corner). A corner is kept iff its score is strictly larger than the scores of its 8 neighbours,
// a neighbour that is not a corner at T (or is outside the 3-px interior) counting as 0. Row-major output.
// Called per cell at src/ORBextractor.cc:810-826 (TBB) / :923-955 (serial).
// ---------------------------------------------------------------------------------------------------------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline int fast_m(const uint8_t* p, int stride) {
  // max over the 16 arcs of max(min_arc d, min_arc -d): windows of 9 are built from windows of 3
  int d[16], lo3[16], hi3[16];
  const int v = p[0];
  for (int k = 0; k < 16; k++) d[k] = v - p[kRingDy[k] * stride + kRingDx[k]];
  for (int k = 0; k < 16; k++) {
    const int a = d[k], b = d[(k + 1) & 15], c = d[(k + 2) & 15];
    lo3[k] = std::min(a, std::min(b, c));
    hi3[k] = std::max(a, std::max(b, c));
  }
  int best = -256;
  for (int s = 0; s < 16; s++) {
    const int lo = std::min(lo3[s], std::min(lo3[(s + 3) & 15], lo3[(s + 6) & 15]));
    const int hi = std::max(hi3[s], std::max(hi3[(s + 3) & 15], hi3[(s + 6) & 15]));
    best = std::max(best, std::max(lo, -hi));
  }
  return best;
}

struct FastPt {
  int x, y, score;
};

inline bool has_arc9(unsigned mask) {
  unsigned m2 = mask | (mask << 16);
  unsigned r = m2 & (m2 >> 1);
  r &= r >> 2;
  r &= r >> 4;
  return (r & (m2 >> 8)) != 0;
}

ORBREF_SIMD_CLONES void fast9_nms(const uint8_t* img, int w, int h, int stride, int threshold, std::vector<FastPt>& out) {
  out.clear();
  threshold = std::min(std::max(threshold, 0), 255);
  if (w < 7 || h < 7) return;
  // Three score rows are enough for the 3x3 non-max suppression (OpenCV keeps a rolling buffer the same way). Each row
  // also keeps the list of its corner columns (ascending): clearing a row and suppressing it then cost as much as the
  // row has corners, not as it has pixels. thread_local: cv::FAST is called per 35 x 35 cell, ~600 times per frame.
  thread_local std::vector<int> rows;
  thread_local std::vector<int> cols[3];
  thread_local std::vector<uint8_t> pass, tmp;
  rows.assign((size_t)3 * w, 0);
  pass.assign((size_t)w + 16, 0);
  tmp.resize((size_t)4 * w);
  for (auto& cvec : cols) cvec.clear();
  int off[16];
  for (int k = 0; k < 16; k++) off[k] = kRingDy[k] * stride + kRingDx[k];
  const int T = threshold;
  auto emit_row = [&](int y) {  // NMS of row y: needs the score rows y - 1, y, y + 1
    const int* up = &rows[(size_t)((y - 1) % 3) * w];
    const int* c = &rows[(size_t)(y % 3) * w];
    const int* dn = &rows[(size_t)((y + 1) % 3) * w];
    for (int x : cols[y % 3]) {
      const int s = c[x];
      if (s == 0) continue;  // OpenCV keeps scores in a buffer where 0 = "not a corner"; a 0-score corner never wins
      if (s > c[x - 1] && s > c[x + 1] && s > up[x - 1] && s > up[x] && s > up[x + 1] && s > dn[x - 1] && s > dn[x] &&
          s > dn[x + 1])
        out.push_back({x, y, s});
    }
  };
  auto clear_row = [&](int slot) {
    int* sc = &rows[(size_t)slot * w];
    for (int x : cols[slot]) sc[x] = 0;
    cols[slot].clear();
  };
  for (int y = 3; y < h - 3; y++) {
    const uint8_t* r = img + (size_t)y * stride;
    int* sc = &rows[(size_t)(y % 3) * w];
    clear_row(y % 3);
    std::vector<int>& mine = cols[y % 3];
    // the exact test + score of a pixel that passed the pre-test
    auto candidate = [&](int x) {
      const uint8_t* p = r + x;
      const int v = p[0];
      unsigned ma = 0, mb = 0;
      for (int k = 0; k < 16; k++) {
        int d = v - p[off[k]];
        ma |= (unsigned)(d > T) << k;
        mb |= (unsigned)(d < -T) << k;
      }
      if (!has_arc9(ma) && !has_arc9(mb)) return;
      sc[x] = fast_m(p, stride) - 1;  // m > T is guaranteed here
      mine.push_back(x);
    };
    // pre-test: every arc of 9 ring pixels contains one end of each of the 8 diameters (k, k + 8): all 8 diameters must
    // have a brighter end, or all 8 a darker end
    const int n = w - 6;
#if defined(__SSE2__)
    if (n >= 16) {
      // 16 pixels per step; a block where no pixel passes the first two diameters (the compass points) is skipped, as
      // OpenCV's own SIMD loop does. The last block is re-aligned to end at n: only its new pixels are taken.
      const __m128i sign = _mm_set1_epi8((char)0x80), tt = _mm_set1_epi8((char)T);
      const uint8_t* c = r + 3;
      static const int order[8] = {0, 4, 2, 6, 1, 3, 5, 7};
      auto block = [&](int i, int first_new) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c + i));
        const __m128i vbs = _mm_xor_si128(_mm_adds_epu8(v, tt), sign), vds = _mm_xor_si128(_mm_subs_epu8(v, tt), sign);
        __m128i brm = _mm_set1_epi8((char)0xff), dkm = brm;
        for (int q = 0; q < 8; q++) {
          const int k = order[q];
          const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c + i + off[k]));
          const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c + i + off[k + 8]));
          const __m128i hi = _mm_xor_si128(_mm_max_epu8(a, b), sign), lo = _mm_xor_si128(_mm_min_epu8(a, b), sign);
          brm = _mm_and_si128(brm, _mm_cmpgt_epi8(hi, vbs));
          dkm = _mm_and_si128(dkm, _mm_cmpgt_epi8(vds, lo));
          if (q == 1 && _mm_movemask_epi8(_mm_or_si128(brm, dkm)) == 0) return;
        }
        unsigned m = (unsigned)_mm_movemask_epi8(_mm_or_si128(brm, dkm));
        m &= ~0u << first_new;
        while (m) {
          const int bit = __builtin_ctz(m);
          m &= m - 1;
          candidate(3 + i + bit);
        }
      };
      int i = 0;
      for (; i + 16 <= n; i += 16) block(i, 0);
      if (i < n) block(n - 16, i - (n - 16));
    } else
#endif
    {
      uint8_t* __restrict__ ps = pass.data();
      uint8_t* __restrict__ vb = tmp.data();
      uint8_t* __restrict__ vd = tmp.data() + w;
      uint8_t* __restrict__ br = tmp.data() + 2 * (size_t)w;
      uint8_t* __restrict__ dk = tmp.data() + 3 * (size_t)w;
      const uint8_t* __restrict__ c = r + 3;
      for (int i = 0; i < n; i++) {
        const int v = c[i];
        vb[i] = (uint8_t)(v + T > 255 ? 255 : v + T);
        vd[i] = (uint8_t)(v - T < 0 ? 0 : v - T);
        br[i] = 1;
        dk[i] = 1;
      }
      for (int k = 0; k < 8; k++) {
        const uint8_t* __restrict__ qa = r + 3 + off[k];
        const uint8_t* __restrict__ qb = r + 3 + off[k + 8];
        for (int i = 0; i < n; i++) {
          const uint8_t a = qa[i], b = qb[i];
          const uint8_t hi = a > b ? a : b, lo = a < b ? a : b;
          br[i] &= (uint8_t)(hi > vb[i]);
          dk[i] &= (uint8_t)(lo < vd[i]);
        }
      }
      for (int i = 0; i < n; i++) ps[i + 3] = br[i] | dk[i];
      for (int x = 3; x < w - 3; x++)
        if (ps[x]) candidate(x);
    }
    if (y >= 5) emit_row(y - 1);  // rows y - 2, y - 1, y are final (row 3's upper neighbour row 2 is all zeros)
    else if (y == 4) emit_row(3);
  }
  // last interior row: its lower neighbour (row h - 3) holds no corners
  if (h - 4 >= 3) {
    clear_row((h - 3) % 3);
    emit_row(h - 4);
  }
}

// cv::fastAtan2 — OpenCV core mathfuncs_core.simd.hpp atan_f32 (degrees). Called at src/ORBextractor.cc:98.
float fast_atan2(float y, float x) {
  static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
  static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
  static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
  static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
  const float eps = (float)2.2204460492503131e-16;  // (float)DBL_EPSILON
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + eps);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + eps);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---------------------------------------------------------------------------------------------------------------
// Quadtree culling: ORBextractor::DistributeOctTree + ExtractorNode::DivideNode + compareNodes
// (src/ORBextractor.cc:557-757, 490-540, 542-555). Keys are indices into the candidate array, so a node's key order
// is the candidates' order (stable partition), exactly as the reference's vector<KeyPoint> copies.
// ---------------------------------------------------------------------------------------------------------------
struct QNode {
  std::vector<int> keys;
  int ulx = 0, uly = 0, urx = 0, ury = 0, blx = 0, bly = 0, brx = 0, bry = 0;
  std::list<QNode>::iterator self;
  bool leaf = false;  // bNoMore
};

void split_node(const QNode& p, const std::vector<orbx_kp>& cand, QNode (&c)[4]) {
  const int halfX = (int)std::ceil(static_cast<float>(p.urx - p.ulx) / 2);
  const int halfY = (int)std::ceil(static_cast<float>(p.bry - p.uly) / 2);
  const int mx = p.ulx + halfX, my = p.uly + halfY;
  // child 0 = upper-left, 1 = upper-right, 2 = lower-left, 3 = lower-right (n1..n4)
  c[0].ulx = p.ulx; c[0].uly = p.uly; c[0].urx = mx;    c[0].ury = p.uly; c[0].blx = p.ulx; c[0].bly = my;    c[0].brx = mx;    c[0].bry = my;
  c[1].ulx = mx;    c[1].uly = p.uly; c[1].urx = p.urx; c[1].ury = p.ury; c[1].blx = mx;    c[1].bly = my;    c[1].brx = p.urx; c[1].bry = my;
  c[2].ulx = p.ulx; c[2].uly = my;    c[2].urx = mx;    c[2].ury = my;    c[2].blx = p.blx; c[2].bly = p.bly; c[2].brx = mx;    c[2].bry = p.bly;
  c[3].ulx = mx;    c[3].uly = my;    c[3].urx = p.urx; c[3].ury = my;    c[3].blx = mx;    c[3].bly = p.bly; c[3].brx = p.brx; c[3].bry = p.bry;
  for (int k : p.keys) {
    const orbx_kp& kp = cand[k];
    int q;
    if (kp.x < (float)c[0].urx) q = (kp.y < (float)c[0].bry) ? 0 : 2;
    else q = (kp.y < (float)c[0].bry) ? 1 : 3;
    c[q].keys.push_back(k);
  }
  for (auto& n : c)
    if (n.keys.size() == 1) n.leaf = true;
}

typedef std::pair<int, QNode*> SizedNode;
bool sized_node_less(SizedNode& a, SizedNode& b) {
  if (a.first < b.first) return true;
  if (a.first > b.first) return false;
  return a.second->ulx < b.second->ulx;
}

std::vector<orbx_kp> distribute_quadtree(const std::vector<orbx_kp>& cand, int minX, int maxX, int minY, int maxY,
                                         int N) {
  const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
  const float hX = static_cast<float>(maxX - minX) / nIni;
  std::list<QNode> nodes;
  std::vector<QNode*> roots(nIni);
  for (int i = 0; i < nIni; i++) {
    QNode n;
    n.ulx = (int)(hX * static_cast<float>(i));  // cv::Point2i(float, int): C++ truncation
    n.uly = 0;
    n.urx = (int)(hX * static_cast<float>(i + 1));
    n.ury = 0;
    n.blx = n.ulx;
    n.bly = maxY - minY;
    n.brx = n.urx;
    n.bry = maxY - minY;
    nodes.push_back(n);
    roots[i] = &nodes.back();
  }
  for (int k = 0; k < (int)cand.size(); k++) roots[(int)(cand[k].x / hX)]->keys.push_back(k);
  for (auto it = nodes.begin(); it != nodes.end();) {
    if (it->keys.size() == 1) { it->leaf = true; ++it; }
    else if (it->keys.empty()) it = nodes.erase(it);
    else ++it;
  }

  std::vector<SizedNode> pending;
  auto add_children = [&](QNode (&c)[4], int* n_expand) {
    for (int q = 0; q < 4; q++) {
      if (c[q].keys.empty()) continue;
      nodes.push_front(c[q]);
      if (c[q].keys.size() > 1) {
        if (n_expand) ++*n_expand;
        pending.push_back(std::make_pair((int)c[q].keys.size(), &nodes.front()));
        nodes.front().self = nodes.begin();
      }
    }
  };

  bool done = false;
  while (!done) {
    int prev = (int)nodes.size();
    int nToExpand = 0;
    pending.clear();
    for (auto it = nodes.begin(); it != nodes.end();) {
      if (it->leaf) { ++it; continue; }
      QNode c[4];
      split_node(*it, cand, c);
      add_children(c, &nToExpand);
      it = nodes.erase(it);
    }
    if ((int)nodes.size() >= N || (int)nodes.size() == prev) {
      done = true;
    } else if ((int)nodes.size() + nToExpand * 3 > N) {
      while (!done) {
        prev = (int)nodes.size();
        std::vector<SizedNode> work = pending;
        pending.clear();
        std::sort(work.begin(), work.end(), sized_node_less);  // unstable: libstdc++ introsort permutation matters
        for (int j = (int)work.size() - 1; j >= 0; j--) {
          QNode c[4];
          split_node(*work[j].second, cand, c);
          add_children(c, nullptr);
          nodes.erase(work[j].second->self);
          if ((int)nodes.size() >= N) break;
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev) done = true;
      }
    }
  }

  std::vector<orbx_kp> result;
  result.reserve(nodes.size());
  for (auto& n : nodes) {
    int best = n.keys[0];
    float maxResp = cand[best].response;
    for (size_t k = 1; k < n.keys.size(); k++)
      if (cand[n.keys[k]].response > maxResp) { best = n.keys[k]; maxResp = cand[best].response; }
    result.push_back(cand[best]);
  }
  return result;
}

// IC_Angle — src/ORBextractor.cc:75-99
float ic_angle(const uint8_t* img, int stride, float px, float py, const int* umax) {
  int m01 = 0, m10 = 0;
  const uint8_t* c = img + (size_t)cv_round(py) * stride + cv_round(px);
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int vsum = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int a = c[u + v * stride], b = c[u - v * stride];
      vsum += a - b;
      m10 += u * (a + b);
    }
    m01 += v * vsum;
  }
  return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor — src/ORBextractor.cc:100-147. cos/sin are glibc cosf/sinf on a float argument.
void orb_descriptor(const orbx_kp& kp, const uint8_t* img, int stride, uint8_t* desc) {
  const float factorPI = (float)(M_PI / 180.f);
  float angle = kp.angle * factorPI;
  float a = cosf(angle), b = sinf(angle);
  const uint8_t* c = img + (size_t)cv_round(kp.y) * stride + cv_round(kp.x);
  for (int i = 0; i < 32; i++) {
    int byte = 0;
    for (int j = 0; j < 8; j++) {
      const int8_t* p = &kPattern[(i * 8 + j) * 4];
      float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
      int t0 = c[cv_round(x0 * b + y0 * a) * stride + cv_round(x0 * a - y0 * b)];
      int t1 = c[cv_round(x1 * b + y1 * a) * stride + cv_round(x1 * a - y1 * b)];
      byte |= (t0 < t1) << j;
    }
    desc[i] = (uint8_t)byte;
  }
}

struct Level {
  int w = 0, h = 0, bstride = 0;
  std::vector<uint8_t> bordered;  // (w + 38) x (h + 38)
  std::vector<uint8_t> blurred;   // w x h
  std::vector<orbx_kp> cand, kps;
  const uint8_t* roi() const { return bordered.data() + (size_t)kEdge * bstride + kEdge; }
};

}  // namespace

struct orbref_extractor {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;  // include/ORBextractor.h:106 — a double member initialised from a float
  std::vector<float> sf, inv_sf, s2, inv_s2;
  std::vector<int> per_level, umax;
  std::vector<Level> lv;
};

extern "C" {

void orbref_resize_linear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
  resize_linear(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orbref_gauss7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  gauss7(src, w, h, sstride, dst, dstride);
}
void orbref_border101(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int border) {
  border101(src, w, h, sstride, dst, dstride, border);
}
int orbref_fast9(const uint8_t* img, int w, int h, int stride, int threshold, int* xs, int* ys, int* scores, int cap) {
  std::vector<FastPt> pts;
  fast9_nms(img, w, h, stride, threshold, pts);
  for (int i = 0; i < (int)pts.size() && i < cap; i++) {
    xs[i] = pts[i].x;
    ys[i] = pts[i].y;
    scores[i] = pts[i].score;
  }
  return (int)pts.size();
}
const char* orbref_primitives_kind() { return ORBREF_PRIMITIVES_KIND; }
float orbref_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int orbref_cv_round(float v) { return cv_round(v); }

void orbref_std_sort_perm(const int* key0, const int* key1, int n, int* perm) {
  std::vector<QNode> store(n);
  std::vector<SizedNode> v(n);
  for (int i = 0; i < n; i++) {
    store[i].ulx = key1[i];
    v[i] = std::make_pair(key0[i], &store[i]);
  }
  std::sort(v.begin(), v.end(), sized_node_less);
  for (int i = 0; i < n; i++) perm[i] = (int)(v[i].second - store.data());
}

// ORBextractor::ORBextractor — src/ORBextractor.cc:408-469
orbref_extractor* orbref_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
  auto* ex = new orbref_extractor;
  ex->nfeatures = nfeatures;
  ex->scaleFactor = scale_factor;
  ex->nlevels = nlevels;
  ex->iniTh = ini_th;
  ex->minTh = min_th;
  ex->sf.resize(nlevels);
  ex->s2.resize(nlevels);
  ex->sf[0] = 1.0f;
  ex->s2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    ex->sf[i] = (float)(ex->sf[i - 1] * ex->scaleFactor);  // float * double -> double -> float
    ex->s2[i] = ex->sf[i] * ex->sf[i];
  }
  ex->inv_sf.resize(nlevels);
  ex->inv_s2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) {
    ex->inv_sf[i] = 1.0f / ex->sf[i];
    ex->inv_s2[i] = 1.0f / ex->s2[i];
  }
  ex->per_level.resize(nlevels);
  float factor = (float)(1.0f / ex->scaleFactor);
  float want = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; l++) {
    ex->per_level[l] = cv_round(want);
    sum += ex->per_level[l];
    want *= factor;
  }
  ex->per_level[nlevels - 1] = std::max(nfeatures - sum, 0);

  ex->umax.assign(kHalfPatch + 1, 0);
  int v, v0, vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
  int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
  const double hp2 = kHalfPatch * kHalfPatch;
  for (v = 0; v <= vmax; ++v) ex->umax[v] = cv_round(std::sqrt(hp2 - v * v));
  for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
    while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0;
    ex->umax[v] = v0;
    ++v0;
  }
  ex->lv.resize(nlevels);
  return ex;
}

void orbref_extractor_destroy(orbref_extractor* ex) { delete ex; }

void orbref_extractor_tables(const orbref_extractor* ex, float* scale, float* inv_scale, float* sigma2,
                             float* inv_sigma2, int* features_per_level, int* umax) {
  for (int i = 0; i < ex->nlevels; i++) {
    if (scale) scale[i] = ex->sf[i];
    if (inv_scale) inv_scale[i] = ex->inv_sf[i];
    if (sigma2) sigma2[i] = ex->s2[i];
    if (inv_sigma2) inv_sigma2[i] = ex->inv_s2[i];
    if (features_per_level) features_per_level[i] = ex->per_level[i];
  }
  if (umax)
    for (int i = 0; i <= kHalfPatch; i++) umax[i] = ex->umax[i];
}

// ORBextractor::operator() — src/ORBextractor.cc:1015-1106, with ComputePyramid (:1108-1145), the SERIAL
// ComputeKeyPointsOctTree (:886-999) and level-ascending assembly.
int orbref_extract(orbref_extractor* ex, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                   orbx_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_index) {
  if (n_out) *n_out = 0;
  if (mono_index) *mono_index = 0;
  if (!img || w <= 0 || h <= 0) return -1;  // :1021
  const int L = ex->nlevels;

  // ---- ComputePyramid ----
  for (int l = 0; l < L; l++) {
    Level& lv = ex->lv[l];
    float s = ex->inv_sf[l];
    lv.w = cv_round((float)w * s);
    lv.h = cv_round((float)h * s);
    lv.bstride = lv.w + 2 * kEdge;
    lv.bordered.assign((size_t)lv.bstride * (lv.h + 2 * kEdge), 0);
    uint8_t* roi = lv.bordered.data() + (size_t)kEdge * lv.bstride + kEdge;
    if (l == 0) {
      border101(img, w, h, stride, lv.bordered.data(), lv.bstride, kEdge);
    } else {
      const Level& pv = ex->lv[l - 1];
      resize_linear(pv.roi(), pv.w, pv.h, pv.bstride, roi, lv.w, lv.h, lv.bstride);
      std::vector<uint8_t> tmp((size_t)lv.w * lv.h);
      for (int y = 0; y < lv.h; y++) memcpy(&tmp[(size_t)y * lv.w], roi + (size_t)y * lv.bstride, lv.w);
      border101(tmp.data(), lv.w, lv.h, lv.w, lv.bordered.data(), lv.bstride, kEdge);
    }
  }

  // ---- ComputeKeyPointsOctTree (serial) ----
  const float W = 35;
  for (int l = 0; l < L; l++) {
    Level& lv = ex->lv[l];
    const uint8_t* roi = lv.roi();
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = lv.w - kEdge + 3, maxBY = lv.h - kEdge + 3;
    lv.cand.clear();
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<FastPt> cell;
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        const uint8_t* sub = roi + (size_t)(int)iniY * lv.bstride + (int)iniX;
        const int cw = (int)maxX - (int)iniX, ch = (int)maxY - (int)iniY;
        fast9_nms(sub, cw, ch, lv.bstride, ex->iniTh, cell);
        if (cell.empty()) fast9_nms(sub, cw, ch, lv.bstride, ex->minTh, cell);
        for (const FastPt& p : cell) {
          orbx_kp kp;
          kp.x = (float)p.x;
          kp.y = (float)p.y;
          kp.x += j * wCell;
          kp.y += i * hCell;
          kp.size = 7.f;
          kp.angle = -1.f;
          kp.response = (float)p.score;
          kp.octave = 0;
          kp.class_id = -1;
          lv.cand.push_back(kp);
        }
      }
    }
    lv.kps = distribute_quadtree(lv.cand, minBX, maxBX, minBY, maxBY, ex->per_level[l]);
    const int scaledPatch = (int)(kPatch * ex->sf[l]);
    for (orbx_kp& kp : lv.kps) {
      kp.x += minBX;
      kp.y += minBY;
      kp.octave = l;
      kp.size = (float)scaledPatch;
    }
  }
  for (int l = 0; l < L; l++) {
    Level& lv = ex->lv[l];
    for (orbx_kp& kp : lv.kps) kp.angle = ic_angle(lv.roi(), lv.bstride, kp.x, kp.y, ex->umax.data());
  }

  // ---- descriptors + assembly ----
  int total = 0;
  for (int l = 0; l < L; l++) total += (int)ex->lv[l].kps.size();
  if (n_out) *n_out = total;
  if (total > cap) return -2;
  int mono = 0, stereo = total - 1;
  for (int l = 0; l < L; l++) {
    Level& lv = ex->lv[l];
    lv.blurred.clear();
    if (lv.kps.empty()) continue;  // :1070 — no blur for empty levels
    std::vector<uint8_t> clone((size_t)lv.w * lv.h);
    for (int y = 0; y < lv.h; y++) memcpy(&clone[(size_t)y * lv.w], lv.roi() + (size_t)y * lv.bstride, lv.w);
    lv.blurred.resize((size_t)lv.w * lv.h);
    gauss7(clone.data(), lv.w, lv.h, lv.w, lv.blurred.data(), lv.w);
    const float scale = ex->sf[l];
    for (const orbx_kp& k0 : lv.kps) {
      uint8_t d[32];
      orb_descriptor(k0, lv.blurred.data(), lv.w, d);
      orbx_kp kp = k0;
      if (l != 0) { kp.x *= scale; kp.y *= scale; }
      int dst;
      if (kp.x >= (float)lap0 && kp.x <= (float)lap1) dst = stereo--;
      else dst = mono++;
      kps[dst] = kp;
      memcpy(desc + (size_t)dst * 32, d, 32);
    }
  }
  if (mono_index) *mono_index = mono;
  return 0;
}

int orbref_level_dims(const orbref_extractor* ex, int level, int* w, int* h) {
  if (level < 0 || level >= ex->nlevels) return -1;
  if (w) *w = ex->lv[level].w;
  if (h) *h = ex->lv[level].h;
  return 0;
}
const uint8_t* orbref_level_image(const orbref_extractor* ex, int level, int* stride) {
  if (stride) *stride = ex->lv[level].bstride;
  return ex->lv[level].roi();
}
const uint8_t* orbref_level_bordered(const orbref_extractor* ex, int level, int* stride) {
  if (stride) *stride = ex->lv[level].bstride;
  return ex->lv[level].bordered.data();
}
const uint8_t* orbref_level_blurred(const orbref_extractor* ex, int level, int* stride) {
  if (stride) *stride = ex->lv[level].w;
  return ex->lv[level].blurred.empty() ? nullptr : ex->lv[level].blurred.data();
}
int orbref_level_candidates(const orbref_extractor* ex, int level, orbx_kp* out, int cap) {
  const auto& v = ex->lv[level].cand;
  for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
  return (int)v.size();
}
int orbref_level_keypoints(const orbref_extractor* ex, int level, orbx_kp* out, int cap) {
  const auto& v = ex->lv[level].kps;
  for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
  return (int)v.size();
}

// ---------------------------------------------------------------------------------------------------------------
// Matching
// ---------------------------------------------------------------------------------------------------------------

// ORBmatcher::DescriptorDistance — src/ORBmatcher.cc:1959-1973 (SWAR popcount over 8 x 32 bit)
int orbref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t x, y;
    memcpy(&x, a + 4 * i, 4);
    memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k = 2) — src/Frame.cc:1293
void orbref_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx1, int32_t* d1, int32_t* idx2,
                 int32_t* d2) {
  for (int i = 0; i < nq; i++) {
    int b1 = INT_MAX, b2 = INT_MAX, i1 = -1, i2 = -1;
    const uint64_t* a = reinterpret_cast<const uint64_t*>(q + (size_t)i * 32);
    uint64_t a0, a1, a2, a3;
    memcpy(&a0, a, 8); memcpy(&a1, a + 1, 8); memcpy(&a2, a + 2, 8); memcpy(&a3, a + 3, 8);
    for (int j = 0; j < nt; j++) {
      uint64_t c0, c1, c2, c3;
      const uint8_t* tp = t + (size_t)j * 32;
      memcpy(&c0, tp, 8); memcpy(&c1, tp + 8, 8); memcpy(&c2, tp + 16, 8); memcpy(&c3, tp + 24, 8);
      int d = __builtin_popcountll(a0 ^ c0) + __builtin_popcountll(a1 ^ c1) + __builtin_popcountll(a2 ^ c2) +
              __builtin_popcountll(a3 ^ c3);
      if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = j; }
      else if (d < b2) { b2 = d; i2 = j; }
    }
    idx1[i] = i1; d1[i] = i1 < 0 ? -1 : b1;
    idx2[i] = i2; d2[i] = i2 < 0 ? -1 : b2;
  }
}

// Frame::ComputeStereoMatches — src/Frame.cc:921-1084
int orbref_stereo_match(const orbref_extractor* left, const orbref_extractor* right, const orbx_kp* kps_l,
                        const uint8_t* desc_l, int n_l, const orbx_kp* kps_r, const uint8_t* desc_r, int n_r,
                        float mbf, float mb, float* u_right, float* depth) {
  const int TH_HIGH = 100, TH_LOW = 50;
  for (int i = 0; i < n_l; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; }
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  const int nRows = left->lv[0].h;
  const std::vector<float>& sf = left->sf;      // Frame::mvScaleFactors = left extractor's (src/Frame.cc:178)
  const std::vector<float>& inv_sf = left->inv_sf;
  std::vector<std::vector<size_t>> rows(nRows);
  for (int iR = 0; iR < n_r; iR++) {
    const orbx_kp& kp = kps_r[iR];
    if (kp.y == 0.0 && kp.x == 0.0) continue;
    const float r = 2.0f * sf[kp.octave];
    const int maxr = (int)std::ceil(kp.y + r);
    const int minr = (int)std::floor(kp.y - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) rows[yi].push_back(iR);  // the reference indexes unchecked; keypoints are >= 19*s inside
  }
  const float minZ = mb, minD = 0, maxD = mbf / minZ;
  std::vector<std::pair<int, int>> distIdx;
  for (int iL = 0; iL < n_l; iL++) {
    const orbx_kp& kpL = kps_l[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const std::vector<size_t>& cands = rows[(size_t)vL];
    if (cands.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH;
    size_t bestIdxR = 0;
    const uint8_t* dL = desc_l + (size_t)iL * 32;
    for (size_t iR : cands) {
      const orbx_kp& kpR = kps_r[iR];
      if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
      const float uR = kpR.x;
      if (uR >= minU && uR <= maxU) {
        int dist = orbref_descriptor_distance(dL, desc_r + iR * 32);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (bestDist < thOrbDist) {
      const float uR0 = kps_r[bestIdxR].x;
      const float scaleFactor = inv_sf[kpL.octave];
      const float scaleduL = std::round(kpL.x * scaleFactor);
      const float scaledvL = std::round(kpL.y * scaleFactor);
      const float scaleduR0 = std::round(uR0 * scaleFactor);
      const int w = 5, L = 5;
      const Level& pl = left->lv[kpL.octave];
      const Level& pr = right->lv[kpL.octave];
      int bestSad = INT_MAX, bestinc = 0;
      float dists[2 * 5 + 1];
      const float iniu = scaleduR0 + L - w;
      const float endu = scaleduR0 + L + w + 1;
      if (iniu < 0 || endu >= pr.w) continue;
      const int yl = (int)(scaledvL - w), xl = (int)(scaleduL - w);
      for (int inc = -L; inc <= L; inc++) {
        const int xr = (int)(scaleduR0 + inc - w);
        double acc = 0;  // cv::norm(IL, IR, NORM_L1)
        for (int yy = 0; yy < 2 * w + 1; yy++) {
          const uint8_t* a = pl.roi() + (ptrdiff_t)(yl + yy) * pl.bstride + xl;
          const uint8_t* b = pr.roi() + (ptrdiff_t)(yl + yy) * pr.bstride + xr;
          for (int xx = 0; xx < 2 * w + 1; xx++) acc += std::abs((int)a[xx] - (int)b[xx]);
        }
        float dist = (float)acc;
        if (dist < (float)bestSad) { bestSad = (int)dist; bestinc = inc; }
        dists[L + inc] = dist;
      }
      if (bestinc == -L || bestinc == L) continue;
      const float dist1 = dists[L + bestinc - 1], dist2 = dists[L + bestinc], dist3 = dists[L + bestinc + 1];
      const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
      if (deltaR < -1 || deltaR > 1) continue;
      float bestuR = sf[kpL.octave] * ((float)scaleduR0 + (float)bestinc + deltaR);
      float disparity = (uL - bestuR);
      if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) {
          disparity = 0.01;
          bestuR = uL - 0.01;
        }
        depth[iL] = mbf / disparity;
        u_right[iL] = bestuR;
        distIdx.push_back(std::make_pair(bestSad, iL));
      }
    }
  }
  if (distIdx.empty()) return 0;  // the reference reads vDistIdx[0] here (UB); defined as "no matches"
  std::sort(distIdx.begin(), distIdx.end());
  const float median = (float)distIdx[distIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = (int)distIdx.size();
  for (int i = (int)distIdx.size() - 1; i >= 0; i--) {
    if (distIdx[i].first < thDist) break;
    u_right[distIdx[i].second] = -1;
    depth[distIdx[i].second] = -1;
    kept--;
  }
  return kept;
}

// Frame::AssignFeaturesToGrid + PosInGrid — src/Frame.cc:520-547, 833-844
void orbref_build_grid(const orbx_kp* kps, int n, float min_x, float min_y, float inv_w, float inv_h,
                       int32_t* offsets, int32_t* items) {
  const int C = ORBX_GRID_COLS, R = ORBX_GRID_ROWS;
  std::vector<std::vector<int>> cells(C * R);
  for (int i = 0; i < n; i++) {
    int px = (int)std::round((kps[i].x - min_x) * inv_w);
    int py = (int)std::round((kps[i].y - min_y) * inv_h);
    if (px < 0 || px >= C || py < 0 || py >= R) continue;
    cells[px * R + py].push_back(i);
  }
  int o = 0;
  for (int c = 0; c < C * R; c++) {
    offsets[c] = o;
    for (int i : cells[c]) items[o++] = i;
  }
  offsets[C * R] = o;
}

// Frame::isInFrustum (src/Frame.cc:632-699, the Nleft == -1 branch) with MapPoint::PredictScale (src/MapPoint.cc:559-573),
// MapPoint::Get{Min,Max}DistanceInvariance (:533-541) and Pinhole::project (src/CameraModels/Pinhole.cpp:47-53), applied
// to the points of one local map the way the loop of Tracking::SearchLocalPoints does (src/Tracking.cc:3288-3300).
// Float expressions are written in the order Eigen evaluates them for fixed-size 3-vectors without packet access: a
// 3-term reduction (dot, squaredNorm, a row of Matrix3f * Vector3f) is c0 + (c1 + c2) (redux_novec_unroller splits
// the range in halves, Eigen/src/Core/Redux.h). log() is the host's glibc logf.
void orbref_is_in_frustum(const orbx_frustum* fr, const orbx_local_map* map, int map_index, float viewing_cos_limit,
                          uint8_t* track_in_view, float* proj_x, float* proj_y, float* proj_xr, int32_t* level,
                          float* view_cos, float* depth) {
  const size_t base = (size_t)map_index * (size_t)map->m;
  auto dot3 = [](const float* a, const float* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); };
  for (int i = 0; i < map->m; i++) {
    const size_t g = base + (size_t)i;
    track_in_view[i] = 0;
    if (map->skip && map->skip[g]) continue;  // Tracking.cc:3289: seen in this frame already, or bad
    proj_x[i] = -1;                            // :635-636
    proj_y[i] = -1;
    const float* P = map->pos + 3 * g;
    const float Pc[3] = {dot3(fr->Rcw + 0, P) + fr->tcw[0], dot3(fr->Rcw + 3, P) + fr->tcw[1],
                         dot3(fr->Rcw + 6, P) + fr->tcw[2]};
    const float Pc_dist = std::sqrt(dot3(Pc, Pc));
    const float PcZ = Pc[2];
    const float invz = 1.0f / PcZ;
    if (PcZ < 0.0f) continue;
    const float u = fr->fx * Pc[0] / Pc[2] + fr->cx;
    const float v = fr->fy * Pc[1] / Pc[2] + fr->cy;
    if (u < fr->min_x || u > fr->max_x) continue;
    if (v < fr->min_y || v > fr->max_y) continue;
    proj_x[i] = u;
    proj_y[i] = v;
    const float maxDistance = 1.2f * map->max_dist[g];
    const float minDistance = 0.8f * map->min_dist[g];
    const float PO[3] = {P[0] - fr->Ow[0], P[1] - fr->Ow[1], P[2] - fr->Ow[2]};
    const float dist = std::sqrt(dot3(PO, PO));
    if (dist < minDistance || dist > maxDistance) continue;
    const float viewCos = dot3(PO, map->normal + 3 * g) / dist;
    if (viewCos < viewing_cos_limit) continue;
    const float ratio = map->max_dist[g] / dist;
    int nScale = (int)std::ceil(std::log(ratio) / fr->log_scale_factor);  // float overloads: logf, ceilf
    if (nScale < 0) nScale = 0;
    else if (nScale >= fr->n_levels) nScale = fr->n_levels - 1;
    track_in_view[i] = 1;
    proj_xr[i] = u - fr->mbf * invz;
    depth[i] = Pc_dist;
    level[i] = nScale;
    view_cos[i] = viewCos;
  }
}

// MapPoint::ComputeDistinctiveDescriptors — src/MapPoint.cc:407-435
int orbref_distinctive_descriptor(const uint8_t* desc, int n) {
  if (n <= 0) return -1;
  const size_t N = (size_t)n;
  std::vector<float> D(N * N);  // float Distances[N][N], :412
  for (size_t i = 0; i < N; i++) {
    D[i * N + i] = 0;
    for (size_t j = i + 1; j < N; j++) {
      const int dij = orbref_descriptor_distance(desc + i * 32, desc + j * 32);
      D[i * N + j] = (float)dij;
      D[j * N + i] = (float)dij;
    }
  }
  int BestMedian = INT_MAX, BestIdx = 0;
  for (size_t i = 0; i < N; i++) {
    std::vector<int> vDists(D.begin() + i * N, D.begin() + (i + 1) * N);  // vector<int>(float*, float*), :427
    std::sort(vDists.begin(), vDists.end());
    const int median = vDists[(size_t)(0.5 * (N - 1))];                  // :429
    if (median < BestMedian) {
      BestMedian = median;
      BestIdx = (int)i;
    }
  }
  return BestIdx;
}

// Frame::GetFeaturesInArea — src/Frame.cc:765-831 (Nleft == -1 branch)
int orbref_features_in_area(const orbx_frame_view* f, float x, float y, float r, int minLevel, int maxLevel,
                            int32_t* out) {
  const int C = ORBX_GRID_COLS, R = ORBX_GRID_ROWS;
  const orbx_grid& g = f->grid;
  int n = 0;
  float factorX = r, factorY = r;
  const int nMinCellX = std::max(0, (int)std::floor((x - g.min_x - factorX) * g.inv_w));
  if (nMinCellX >= C) return 0;
  const int nMaxCellX = std::min(C - 1, (int)std::ceil((x - g.min_x + factorX) * g.inv_w));
  if (nMaxCellX < 0) return 0;
  const int nMinCellY = std::max(0, (int)std::floor((y - g.min_y - factorY) * g.inv_h));
  if (nMinCellY >= R) return 0;
  const int nMaxCellY = std::min(R - 1, (int)std::ceil((y - g.min_y + factorY) * g.inv_h));
  if (nMaxCellY < 0) return 0;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      const int c = ix * R + iy;
      for (int j = g.cell_offsets[c]; j < g.cell_offsets[c + 1]; j++) {
        const int idx = g.cell_items[j];
        const orbx_kp& kp = f->kps[idx];
        if (bCheckLevels) {
          if (kp.octave < minLevel) continue;
          if (maxLevel >= 0 && kp.octave > maxLevel) continue;
        }
        const float distx = kp.x - x, disty = kp.y - y;
        if (std::fabs(distx) < factorX && std::fabs(disty) < factorY) out[n++] = idx;
      }
    }
  return n;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, ...) — src/ORBmatcher.cc:42-221, serial order
int orbref_search_by_projection_map(const orbx_frame_view* f, const orbx_mappoints* mps, float th, float nnratio,
                                    int far_points, float th_far, int32_t* assign) {
  const int TH_HIGH = 100;
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  std::vector<uint8_t> occ(f->occupied, f->occupied + f->n);
  std::vector<int32_t> idxs(f->n);
  for (int i = 0; i < f->n; i++) assign[i] = -1;
  for (int iMP = 0; iMP < mps->m; iMP++) {
    if (!mps->track_in_view[iMP]) continue;
    if (far_points && mps->depth[iMP] > th_far) continue;
    const int level = mps->level[iMP];
    float r = ((double)mps->view_cos[iMP] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos :223-228
    if (bFactor) r *= th;
    const float rs = r * f->scale_factors[level];
    const int nc = orbref_features_in_area(f, mps->proj_x[iMP], mps->proj_y[iMP], rs, level - 1, level, idxs.data());
    if (nc == 0) continue;
    const uint8_t* dMP = mps->desc + (size_t)iMP * 32;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int c = 0; c < nc; c++) {
      const int idx = idxs[c];
      if (occ[idx]) continue;
      if (f->u_right && f->u_right[idx] > 0) {
        const float er = std::fabs(mps->proj_xr[iMP] - f->u_right[idx]);
        if (er > rs) continue;
      }
      const int dist = orbref_descriptor_distance(dMP, f->desc + (size_t)idx * 32);
      if (dist < bestDist) {
        bestDist2 = bestDist;
        bestDist = dist;
        bestLevel2 = bestLevel;
        bestLevel = f->kps[idx].octave;
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = f->kps[idx].octave;
        bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
      assign[bestIdx] = iMP;
      occ[bestIdx] = mps->has_obs[iMP];
      nmatches++;
    }
  }
  return nmatches;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, ...) on a two-camera Frame — src/ORBmatcher.cc:42-221
// with Nleft != -1, serial order
int orbref_search_by_projection_map_fisheye(const orbx_fisheye_view* f, const orbx_mappoints* mps,
                                            const orbx_mappoints_right* mr, float th, float nnratio, int far_points,
                                            float th_far, int32_t* assign) {
  const int TH_HIGH = 100;
  const int NL = f->n_left, NR = f->n_right, N = NL + NR;
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  // slot state = F.mvpMapPoints: who sits there (-2 = the frame's own earlier occupant) and whether it has observations
  std::vector<int32_t> who(N, -1);
  std::vector<uint8_t> blocked(f->occupied, f->occupied + N);
  // Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel, bRight) with Nleft != -1 (src/Frame.cc:765-831)
  auto in_area = [&](float x, float y, float r, int minLevel, int maxLevel, bool right, std::vector<int>& out) {
    out.clear();
    const orbx_grid& g = right ? f->grid_right : f->grid_left;
    const orbx_kp* kps = right ? f->kps_right : f->kps_left;
    const float min_x = f->grid_left.min_x, min_y = f->grid_left.min_y, inv_w = f->grid_left.inv_w, inv_h = f->grid_left.inv_h;
    const int C = ORBX_GRID_COLS, R = ORBX_GRID_ROWS;
    const int x0 = std::max(0, (int)std::floor((x - min_x - r) * inv_w));
    if (x0 >= C) return;
    const int x1 = std::min(C - 1, (int)std::ceil((x - min_x + r) * inv_w));
    if (x1 < 0) return;
    const int y0 = std::max(0, (int)std::floor((y - min_y - r) * inv_h));
    if (y0 >= R) return;
    const int y1 = std::min(R - 1, (int)std::ceil((y - min_y + r) * inv_h));
    if (y1 < 0) return;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = x0; ix <= x1; ix++)
      for (int iy = y0; iy <= y1; iy++) {
        const int c = ix * R + iy;
        for (int j = g.cell_offsets[c]; j < g.cell_offsets[c + 1]; j++) {
          const orbx_kp& kp = kps[g.cell_items[j]];
          if (check) {
            if (kp.octave < minLevel) continue;
            if (maxLevel >= 0 && kp.octave > maxLevel) continue;
          }
          if (std::fabs(kp.x - x) < r && std::fabs(kp.y - y) < r) out.push_back(g.cell_items[j]);
        }
      }
  };
  for (int i = 0; i < N; i++) assign[i] = -1;
  std::vector<int> idxs;
  for (int iMP = 0; iMP < mps->m; iMP++) {
    const bool inL = mps->track_in_view[iMP] != 0, inR = mr->track_in_view_r[iMP] != 0;
    if (!inL && !inR) continue;                                     // :55 (the isBad() test of :59 is folded into both flags)
    if (far_points && mps->depth[iMP] > th_far) continue;           // :57
    const uint8_t* dMP = mps->desc + (size_t)iMP * 32;
    const uint8_t obs = mps->has_obs[iMP];
    auto put = [&](int slot) {
      who[slot] = iMP;
      blocked[slot] = obs;
    };
    if (inL) {                                                      // :61-145
      const int level = mps->level[iMP];
      float r = ((double)mps->view_cos[iMP] > 0.998) ? 2.5f : 4.0f;
      if (bFactor) r *= th;
      in_area(mps->proj_x[iMP], mps->proj_y[iMP], r * f->scale_factors[level], level - 1, level, false, idxs);
      if (!idxs.empty()) {
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : idxs) {
          if (blocked[idx]) continue;
          const int dist = orbref_descriptor_distance(dMP, f->desc + (size_t)idx * 32);
          if (dist < bestDist) {
            bestDist2 = bestDist;
            bestDist = dist;
            bestLevel2 = bestLevel;
            bestLevel = f->kps_left[idx].octave;
            bestIdx = idx;
          } else if (dist < bestDist2) {
            bestLevel2 = f->kps_left[idx].octave;
            bestDist2 = dist;
          }
        }
        if (bestDist <= TH_HIGH) {
          // the reference `continue`s here (:125-126): the right-camera search of this point is skipped as well
          if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
          put(bestIdx);
          if (f->left_to_right[bestIdx] != -1) {                    // :131-137
            put(f->left_to_right[bestIdx] + NL);
            nmatches++;
          }
          nmatches++;
        }
      }
    }
    if (inR) {                                                      // :148-217
      const int level = mr->level_r[iMP];
      if (level != -1) {
        const float r = ((double)mr->view_cos_r[iMP] > 0.998) ? 2.5f : 4.0f;  // no th factor here
        in_area(mr->proj_x_r[iMP], mr->proj_y_r[iMP], r * f->scale_factors[level], level - 1, level, true, idxs);
        if (idxs.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : idxs) {
          if (blocked[idx + NL]) continue;
          const int dist = orbref_descriptor_distance(dMP, f->desc + (size_t)(idx + NL) * 32);
          if (dist < bestDist) {
            bestDist2 = bestDist;
            bestDist = dist;
            bestLevel2 = bestLevel;
            bestLevel = f->kps_right[idx].octave;
            bestIdx = idx;
          } else if (dist < bestDist2) {
            bestLevel2 = f->kps_right[idx].octave;
            bestDist2 = dist;
          }
        }
        if (bestDist <= TH_HIGH) {
          if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
          if (f->right_to_left[bestIdx] != -1) {                    // :203-208
            put(f->right_to_left[bestIdx]);
            nmatches++;
          }
          put(bestIdx + NL);
          nmatches++;
        }
      }
    }
  }
  for (int i = 0; i < N; i++) assign[i] = who[i];
  return nmatches;
}

namespace {
// ORBmatcher::ComputeThreeMaxima — src/ORBmatcher.cc:1920-1955
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}
const int kHisto = 30;  // HISTO_LENGTH
inline int rot_bin(float a1, float a2) {
  const float factor = 1.0f / kHisto;
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);
  if (bin == kHisto) bin = 0;
  return bin;
}
}  // namespace

// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) — src/ORBmatcher.cc:1594-1806 (Nleft == -1) and
// (Frame&, KeyFrame*, set, th, ORBdist) — :1808-1918, after the caller-side projection.
int orbref_search_by_projection_frame(const orbx_frame_view* f, const orbx_projected* pts, int max_dist,
                                      int check_orientation, int32_t* assign) {
  int nmatches = 0;
  std::vector<int> rotHist[kHisto];
  std::vector<uint8_t> occ(f->occupied, f->occupied + f->n);
  std::vector<int32_t> idxs(f->n);
  for (int i = 0; i < f->n; i++) assign[i] = -1;
  for (int i = 0; i < pts->m; i++) {
    const int nc = orbref_features_in_area(f, pts->u[i], pts->v[i], pts->radius[i], pts->min_level[i],
                                           pts->max_level[i], idxs.data());
    if (nc == 0) continue;
    const uint8_t* dMP = pts->desc + (size_t)i * 32;
    int bestDist = 256, bestIdx2 = -1;
    for (int c = 0; c < nc; c++) {
      const int i2 = idxs[c];
      if (occ[i2]) continue;
      if (f->u_right && pts->u_right && f->u_right[i2] > 0) {
        const float er = std::fabs(pts->u_right[i] - f->u_right[i2]);
        if (er > pts->radius[i]) continue;
      }
      const int dist = orbref_descriptor_distance(dMP, f->desc + (size_t)i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= max_dist) {
      // bestIdx2 can be -1 only if bestDist stayed 256 > max_dist, so it is valid here
      assign[bestIdx2] = i;
      occ[bestIdx2] = pts->has_obs[i];
      nmatches++;
      if (check_orientation) rotHist[rot_bin(pts->angle[i], f->kps[bestIdx2].angle)].push_back(bestIdx2);
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, kHisto, ind1, ind2, ind3);
    for (int b = 0; b < kHisto; b++) {
      if (b == ind1 || b == ind2 || b == ind3) continue;
      for (int idx : rotHist[b]) {
        assign[idx] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

// The candidate loop of SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, ...) for ONE camera of a
// two-camera frame (src/ORBmatcher.cc:1649-1690 on mvKeys / mGrid, :1711-1755 on mvKeysRight / mGridRight), the
// rotation histogram left to the caller because both cameras share it (:1693-1706, :1757-1778, :1785-1803).
// decisions[i] = the keypoint point i is written to (-1: none), window[i] = |GetFeaturesInArea(...)| of its window (the
// reference `continue`s past the right-camera search when the LEFT window is empty, :1655). Returns the acceptances.
int orbref_search_by_projection_frame_decisions(const orbx_frame_view* f, const orbx_projected* pts, int max_dist,
                                                int32_t* decisions, int32_t* window) {
  int accepted = 0;
  std::vector<uint8_t> occ(f->occupied, f->occupied + f->n);
  std::vector<int32_t> idxs(f->n);
  for (int i = 0; i < pts->m; i++) {
    decisions[i] = -1;
    const int nc = orbref_features_in_area(f, pts->u[i], pts->v[i], pts->radius[i], pts->min_level[i],
                                           pts->max_level[i], idxs.data());
    if (window) window[i] = nc;
    const uint8_t* dMP = pts->desc + (size_t)i * 32;
    int bestDist = 256, bestIdx2 = -1;
    for (int c = 0; c < nc; c++) {
      const int i2 = idxs[c];
      if (occ[i2]) continue;
      if (f->u_right && pts->u_right && f->u_right[i2] > 0) {
        const float er = std::fabs(pts->u_right[i] - f->u_right[i2]);
        if (er > pts->radius[i]) continue;
      }
      const int dist = orbref_descriptor_distance(dMP, f->desc + (size_t)i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= max_dist) {
      decisions[i] = bestIdx2;
      occ[bestIdx2] = pts->has_obs[i];  // a later point is blocked only by a MapPoint with observations (:1663-1665)
      accepted++;
    }
  }
  return accepted;
}

// ORBmatcher::SearchForTriangulation — src/ORBmatcher.cc:886-1106; Pinhole::epipolarConstrain —
// src/CameraModels/Pinhole.cpp:122-149 (F12 supplied by the caller, row-major).
// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) — src/ORBmatcher.cc:230-404, Nleft == -1
int orbref_search_by_bow(const orbx_keyframe_view* kf, const orbx_keyframe_view* frame, float nnratio,
                         int check_orientation, int32_t* matches_f) {
  const int TH_LOW = 50;
  int nmatches = 0;
  std::vector<int> rotHist[kHisto];
  for (int i = 0; i < frame->n; i++) matches_f[i] = -1;  // vpMapPointMatches = vector<MapPoint*>(F.N, NULL)  :235
  const orbx_featvec& vK = kf->featvec;
  const orbx_featvec& vF = frame->featvec;
  int a = 0, b = 0;
  while (a < vK.n_nodes && b < vF.n_nodes) {
    if (vK.node_ids[a] == vF.node_ids[b]) {
      for (int pK = vK.offsets[a]; pK < vK.offsets[a + 1]; pK++) {
        const int realIdxKF = (int)vK.indices[pK];
        if (!kf->has_mappoint[realIdxKF]) continue;  // !pMP || pMP->isBad()                                  :262-264
        const uint8_t* dKF = kf->desc + (size_t)realIdxKF * 32;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int pF = vF.offsets[b]; pF < vF.offsets[b + 1]; pF++) {
          const int realIdxF = (int)vF.indices[pF];
          if (matches_f[realIdxF] >= 0) continue;  //                                                          :280
          const int dist = orbref_descriptor_distance(dKF, frame->desc + (size_t)realIdxF * 32);
          if (dist < bestDist1) {
            bestDist2 = bestDist1;
            bestDist1 = dist;
            bestIdxF = realIdxF;
          } else if (dist < bestDist2) {
            bestDist2 = dist;
          }
        }
        if (bestDist1 <= TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {  //                        :319-321
          matches_f[bestIdxF] = realIdxKF;
          if (check_orientation)
            rotHist[rot_bin(kf->kps[realIdxKF].angle, frame->kps[bestIdxF].angle)].push_back(bestIdxF);
          nmatches++;
        }
      }
      a++;
      b++;
    } else if (vK.node_ids[a] < vF.node_ids[b]) {
      while (a < vK.n_nodes && vK.node_ids[a] < vF.node_ids[b]) a++;  // lower_bound
    } else {
      while (b < vF.n_nodes && vF.node_ids[b] < vK.node_ids[a]) b++;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, kHisto, ind1, ind2, ind3);
    for (int i = 0; i < kHisto; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idxF : rotHist[i]) {
        matches_f[idxF] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

// The same function on a two-camera Frame (F.Nleft != -1, src/ORBmatcher.cc:274-317, 319-365): per KeyFrame feature the
// best two distances are kept separately over the frame's LEFT rows (realIdxF < Nleft) and RIGHT rows; the left best is
// accepted by the ratio test, and — nested inside "bestDist1 <= TH_LOW" (:319) — the right best whenever
// bestDist1R <= TH_LOW (its ratio test is disabled by "|| true", :347-350). kps of both views = left rows then right rows
// (the keypoint selection of :323-335 / :352-362 flattened by the caller).
int orbref_search_by_bow_fisheye(const orbx_keyframe_view* kf, const orbx_keyframe_view* frame, int n_left_f,
                                 float nnratio, int check_orientation, int32_t* matches_f) {
  const int TH_LOW = 50;
  int nmatches = 0;
  std::vector<int> rotHist[kHisto];
  for (int i = 0; i < frame->n; i++) matches_f[i] = -1;
  const orbx_featvec& vK = kf->featvec;
  const orbx_featvec& vF = frame->featvec;
  int a = 0, b = 0;
  while (a < vK.n_nodes && b < vF.n_nodes) {
    if (vK.node_ids[a] == vF.node_ids[b]) {
      for (int pK = vK.offsets[a]; pK < vK.offsets[a + 1]; pK++) {
        const int realIdxKF = (int)vK.indices[pK];
        if (!kf->has_mappoint[realIdxKF]) continue;
        const uint8_t* dKF = kf->desc + (size_t)realIdxKF * 32;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        int bestDist1R = 256, bestIdxFR = -1, bestDist2R = 256;
        for (int pF = vF.offsets[b]; pF < vF.offsets[b + 1]; pF++) {
          const int realIdxF = (int)vF.indices[pF];
          if (matches_f[realIdxF] >= 0) continue;
          const int dist = orbref_descriptor_distance(dKF, frame->desc + (size_t)realIdxF * 32);
          if (realIdxF < n_left_f && dist < bestDist1) {
            bestDist2 = bestDist1;
            bestDist1 = dist;
            bestIdxF = realIdxF;
          } else if (realIdxF < n_left_f && dist < bestDist2) {
            bestDist2 = dist;
          }
          if (realIdxF >= n_left_f && dist < bestDist1R) {
            bestDist2R = bestDist1R;
            bestDist1R = dist;
            bestIdxFR = realIdxF;
          } else if (realIdxF >= n_left_f && dist < bestDist2R) {
            bestDist2R = dist;
          }
        }
        if (bestDist1 <= TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            matches_f[bestIdxF] = realIdxKF;
            if (check_orientation)
              rotHist[rot_bin(kf->kps[realIdxKF].angle, frame->kps[bestIdxF].angle)].push_back(bestIdxF);
            nmatches++;
          }
          if (bestDist1R <= TH_LOW) {
            matches_f[bestIdxFR] = realIdxKF;
            if (check_orientation)
              rotHist[rot_bin(kf->kps[realIdxKF].angle, frame->kps[bestIdxFR].angle)].push_back(bestIdxFR);
            nmatches++;
          }
        }
      }
      a++;
      b++;
    } else if (vK.node_ids[a] < vF.node_ids[b]) {
      while (a < vK.n_nodes && vK.node_ids[a] < vF.node_ids[b]) a++;
    } else {
      while (b < vF.n_nodes && vF.node_ids[b] < vK.node_ids[a]) b++;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, kHisto, ind1, ind2, ind3);
    for (int i = 0; i < kHisto; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idxF : rotHist[i]) {
        matches_f[idxF] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

// cv::remap(..., INTER_LINEAR), CV_8UC1 source, CV_32FC1 maps, BORDER_CONSTANT(0) (OpenCV imgproc/imgwarp.cpp:
// remapBilinear with INTER_BITS = 5, INTER_REMAP_COEF_BITS = 15; the 32 x 32 weight table is exact: w = a * b * 32)
void orbref_remap_linear(const uint8_t* src, int sw, int sh, int sstride, const float* mapx, const float* mapy, int dw,
                         int dh, uint8_t* dst, int dstride) {
  auto px = [&](int y, int x) -> int {
    return (x >= 0 && x < sw && y >= 0 && y < sh) ? src[(size_t)y * sstride + x] : 0;
  };
  for (int y = 0; y < dh; y++)
    for (int x = 0; x < dw; x++) {
      const int sx = (int)lrintf(mapx[(size_t)y * dw + x] * 32.f), sy = (int)lrintf(mapy[(size_t)y * dw + x] * 32.f);
      const int ix = sx >> 5, iy = sy >> 5, fx = sx & 31, fy = sy & 31;
      const int top = (32 - fx) * px(iy, ix) + fx * px(iy, ix + 1);
      const int bot = (32 - fx) * px(iy + 1, ix) + fx * px(iy + 1, ix + 1);
      dst[(size_t)y * dstride + x] = (uint8_t)(((32 - fy) * top + fy * bot + 512) >> 10);
    }
}

// cv::cvtColor(..., COLOR_*2GRAY), 8-bit (OpenCV imgproc/color_rgb: RGB2Gray<uchar>, 15-bit coefficients)
void orbref_cvt_gray(const uint8_t* src, int w, int h, int stride, int channels, int rgb, uint8_t* dst, int dstride) {
  const int BY = 3735, GY = 19235, RY = 9798;
  for (int y = 0; y < h; y++) {
    const uint8_t* s = src + (size_t)y * stride;
    uint8_t* d = dst + (size_t)y * dstride;
    for (int x = 0; x < w; x++, s += channels) {
      const int b = rgb ? s[2] : s[0], g = s[1], r = rgb ? s[0] : s[2];
      d[x] = (uint8_t)((b * BY + g * GY + r * RY + (1 << 14)) >> 15);
    }
  }
}

// TemplatedVocabulary::transform(const TDescriptor&, WordId&, WordValue&, NodeId*, int) — TemplatedVocabulary.h:1218-1262
void orbref_bow_transform(const orbx_vocabulary* voc, const uint8_t* desc, int n, int levelsup, uint32_t* word_id,
                          double* weight, uint32_t* node_id) {
  const int nid_level = voc->depth - levelsup;
  for (int i = 0; i < n; i++) {
    const uint8_t* f = desc + (size_t)i * 32;
    uint32_t nid = 0, final_id = 0;
    int current_level = 0;
    do {
      ++current_level;
      const int c0 = voc->child_offsets[final_id], c1 = voc->child_offsets[final_id + 1];
      final_id = voc->children[c0];
      double best_d = orbref_descriptor_distance(f, voc->descriptors + (size_t)final_id * 32);
      for (int c = c0 + 1; c < c1; c++) {
        const uint32_t id = voc->children[c];
        const double d = orbref_descriptor_distance(f, voc->descriptors + (size_t)id * 32);
        if (d < best_d) {
          best_d = d;
          final_id = id;
        }
      }
      if (current_level == nid_level) nid = final_id;
    } while (voc->child_offsets[final_id + 1] > voc->child_offsets[final_id]);  // !isLeaf()
    word_id[i] = voc->word_id[final_id];
    weight[i] = voc->weight[final_id];
    node_id[i] = nid;
  }
}

// ORBmatcher::SearchForInitialization — src/ORBmatcher.cc:618-764, serial order of the loop body
int orbref_search_for_initialization(const orbx_frame_view* f1, const orbx_frame_view* f2, const float* prev_xy,
                                     int window_size, float nnratio, int check_orientation, int32_t* matches12) {
  const int TH_LOW = 50;
  int nmatches = 0;
  for (int i = 0; i < f1->n; i++) matches12[i] = -1;
  std::vector<int> rotHist[kHisto];
  std::vector<int> vMatchedDistance(std::max(f2->n, 1), INT_MAX), vnMatches21(std::max(f2->n, 1), -1);
  std::vector<int32_t> idxs(std::max(f2->n, 1));
  for (int i1 = 0; i1 < f1->n; i1++) {
    const orbx_kp& kp1 = f1->kps[i1];
    const int level1 = kp1.octave;
    if (level1 > 0) continue;  //                                                                             :642
    const int nc = orbref_features_in_area(f2, prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window_size, level1, level1,
                                           idxs.data());
    if (nc == 0) continue;
    const uint8_t* d1 = f1->desc + (size_t)i1 * 32;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (int c = 0; c < nc; c++) {
      const int i2 = idxs[c];
      const int dist = orbref_descriptor_distance(d1, f2->desc + (size_t)i2 * 32);
      if (vMatchedDistance[i2] <= dist) continue;  //                                                          :685
      if (dist < bestDist) {
        bestDist2 = bestDist;
        bestDist = dist;
        bestIdx2 = i2;
      } else if (dist < bestDist2) {
        bestDist2 = dist;
      }
    }
    if (bestDist <= TH_LOW) {
      if (bestDist < (float)bestDist2 * nnratio) {  //                                                         :700
        if (vnMatches21[bestIdx2] >= 0) {
          matches12[vnMatches21[bestIdx2]] = -1;
          nmatches--;
        }
        matches12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        nmatches++;
        if (check_orientation) rotHist[rot_bin(f1->kps[i1].angle, f2->kps[bestIdx2].angle)].push_back(i1);
      }
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, kHisto, ind1, ind2, ind3);
    for (int i = 0; i < kHisto; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (matches12[idx1] >= 0) {  //                                                                       :748-751
          matches12[idx1] = -1;
          nmatches--;
        }
    }
  }
  return nmatches;
}

// ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight) — the matching loop, src/ORBmatcher.cc:1194-1257
void orbref_fuse_match(const orbx_frame_view* kf, const float* inv_level_sigma2, const orbx_projected* pts,
                       int chi2_gate, int32_t* best_idx, int32_t* best_dist) {
  std::vector<int32_t> idxs(std::max(kf->n, 1));
  for (int i = 0; i < pts->m; i++) {
    const float u = pts->u[i], v = pts->v[i], ur = pts->u_right ? pts->u_right[i] : 0.f;
    const int nPredictedLevel = pts->max_level[i];
    // KeyFrame::GetFeaturesInArea has no level filter (src/KeyFrame.cc:705-749): minLevel = -1, maxLevel = -1
    const int nc = orbref_features_in_area(kf, u, v, pts->radius[i], -1, -1, idxs.data());
    const uint8_t* dMP = pts->desc + (size_t)i * 32;
    int bestDist = 256, bestIdx = -1;
    for (int c = 0; c < nc; c++) {
      const int idx = idxs[c];
      const orbx_kp& kp = kf->kps[idx];
      const int kpLevel = kp.octave;
      if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;  //                          :1221
      if (!chi2_gate) {  // Fuse(KeyFrame*, Sim3f&, ...) has no reprojection gate                               :1363-1367
      } else if (kf->u_right && kf->u_right[idx] >= 0) {  // check reprojection error in stereo                       :1223-1233
        const float ex = u - kp.x, ey = v - kp.y, er = ur - kf->u_right[idx];
        const float e2 = ex * ex + ey * ey + er * er;
        if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
      } else {  //                                                                                            :1235-1243
        const float ex = u - kp.x, ey = v - kp.y;
        const float e2 = ex * ex + ey * ey;
        if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
      }
      const int dist = orbref_descriptor_distance(dMP, kf->desc + (size_t)idx * 32);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx = idx;
      }
    }
    best_idx[i] = bestIdx;
    best_dist[i] = bestDist;
  }
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) — src/ORBmatcher.cc:766-884, NLeft == -1
int orbref_search_by_bow_kf(const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, float nnratio,
                            int check_orientation, int32_t* matches12) {
  const int TH_LOW = 50;
  int nmatches = 0;
  std::vector<int> rotHist[kHisto];
  for (int i = 0; i < kf1->n; i++) matches12[i] = -1;
  std::vector<bool> vbMatched2(std::max(kf2->n, 1), false);
  const orbx_featvec& v1 = kf1->featvec;
  const orbx_featvec& v2 = kf2->featvec;
  int a = 0, b = 0;
  while (a < v1.n_nodes && b < v2.n_nodes) {
    if (v1.node_ids[a] == v2.node_ids[b]) {
      for (int p1 = v1.offsets[a]; p1 < v1.offsets[a + 1]; p1++) {
        const int idx1 = (int)v1.indices[p1];
        if (!kf1->has_mappoint[idx1]) continue;  //                                                           :802-804
        const uint8_t* d1 = kf1->desc + (size_t)idx1 * 32;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int p2 = v2.offsets[b]; p2 < v2.offsets[b + 1]; p2++) {
          const int idx2 = (int)v2.indices[p2];
          if (vbMatched2[idx2] || !kf2->has_mappoint[idx2]) continue;  //                                     :821-823
          const int dist = orbref_descriptor_distance(d1, kf2->desc + (size_t)idx2 * 32);
          if (dist < bestDist1) {
            bestDist2 = bestDist1;
            bestDist1 = dist;
            bestIdx2 = idx2;
          } else if (dist < bestDist2) {
            bestDist2 = dist;
          }
        }
        if (bestDist1 < TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {  //                         :838-840
          matches12[idx1] = bestIdx2;
          vbMatched2[bestIdx2] = true;
          if (check_orientation) rotHist[rot_bin(kf1->kps[idx1].angle, kf2->kps[bestIdx2].angle)].push_back(idx1);
          nmatches++;
        }
      }
      a++;
      b++;
    } else if (v1.node_ids[a] < v2.node_ids[b]) {
      while (a < v1.n_nodes && v1.node_ids[a] < v2.node_ids[b]) a++;  // lower_bound
    } else {
      while (b < v2.n_nodes && v2.node_ids[b] < v1.node_ids[a]) b++;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, kHisto, ind1, ind2, ind3);
    for (int i = 0; i < kHisto; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) {
        matches12[idx1] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

int orbref_search_for_triangulation(const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, const float* F12,
                                    float ep_x, float ep_y, int only_stereo, int coarse, int check_orientation,
                                    int32_t* matches12) {
  const int TH_LOW = 50;
  int nmatches = 0;
  std::vector<int> rotHist[kHisto];
  for (int i = 0; i < kf1->n; i++) matches12[i] = -1;
  const orbx_featvec& v1 = kf1->featvec;
  const orbx_featvec& v2 = kf2->featvec;
  int a = 0, b = 0;
  while (a < v1.n_nodes && b < v2.n_nodes) {
    if (v1.node_ids[a] == v2.node_ids[b]) {
      for (int p1 = v1.offsets[a]; p1 < v1.offsets[a + 1]; p1++) {
        const int idx1 = (int)v1.indices[p1];
        if (kf1->has_mappoint[idx1]) continue;
        const bool bStereo1 = kf1->u_right && kf1->u_right[idx1] >= 0;
        if (only_stereo && !bStereo1) continue;
        const orbx_kp& kp1 = kf1->kps[idx1];
        const uint8_t* d1 = kf1->desc + (size_t)idx1 * 32;
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int p2 = v2.offsets[b]; p2 < v2.offsets[b + 1]; p2++) {
          const int idx2 = (int)v2.indices[p2];
          if (kf2->has_mappoint[idx2]) continue;  // vbMatched2 is never set in the reference
          const bool bStereo2 = kf2->u_right && kf2->u_right[idx2] >= 0;
          if (only_stereo && !bStereo2) continue;
          const int dist = orbref_descriptor_distance(d1, kf2->desc + (size_t)idx2 * 32);
          if (dist > TH_LOW || dist > bestDist) continue;
          const orbx_kp& kp2 = kf2->kps[idx2];
          if (!bStereo1 && !bStereo2) {
            const float distex = ep_x - kp2.x, distey = ep_y - kp2.y;
            if (distex * distex + distey * distey < 100 * kf2->scale_factors[kp2.octave]) continue;
          }
          bool ok = coarse != 0;
          if (!ok) {
            const float ea = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
            const float eb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
            const float ec = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
            const float num = ea * kp2.x + eb * kp2.y + ec;
            const float den = ea * ea + eb * eb;
            if (den != 0) {
              const float dsqr = num * num / den;
              ok = (double)dsqr < 3.84 * (double)kf2->level_sigma2[kp2.octave];
            }
          }
          if (ok) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
          matches12[idx1] = bestIdx2;
          nmatches++;
          if (check_orientation) rotHist[rot_bin(kp1.angle, kf2->kps[bestIdx2].angle)].push_back(idx1);
        }
      }
      a++;
      b++;
    } else if (v1.node_ids[a] < v2.node_ids[b]) {
      while (a < v1.n_nodes && v1.node_ids[a] < v2.node_ids[b]) a++;  // lower_bound
    } else {
      while (b < v2.n_nodes && v2.node_ids[b] < v1.node_ids[a]) b++;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, kHisto, ind1, ind2, ind3);
    for (int i = 0; i < kHisto; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) {
        matches12[idx1] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

// SearchForTriangulation on two-camera KeyFrames (both mpCamera2 set; src/ORBmatcher.cc:886-1106 with :960, :966-971,
// :991-994, :1007-1043): no feature is "stereo" (bStereo1 / bStereo2 are false, so bOnlyStereo matches nothing), the epipole
// gate is skipped (:996), and the epipolar test runs on the camera pair picked by which side of NLeft the two features lie
// on — here in the Pinhole form on F12[pair], pair = 2 * right1 + right2 (the camera models themselves are not restated;
// the product asks the caller's camera objects, see orbref_triangulation_candidates). kps: left rows then right rows.
int orbref_search_for_triangulation_fisheye(const orbx_keyframe_view* kf1, int n_left1, const orbx_keyframe_view* kf2,
                                            int n_left2, const float* F12x4, int only_stereo, int coarse,
                                            int check_orientation, int32_t* matches12) {
  const int TH_LOW = 50;
  int nmatches = 0;
  std::vector<int> rotHist[kHisto];
  for (int i = 0; i < kf1->n; i++) matches12[i] = -1;
  const orbx_featvec& v1 = kf1->featvec;
  const orbx_featvec& v2 = kf2->featvec;
  int a = 0, b = 0;
  while (a < v1.n_nodes && b < v2.n_nodes) {
    if (v1.node_ids[a] == v2.node_ids[b]) {
      for (int p1 = v1.offsets[a]; p1 < v1.offsets[a + 1]; p1++) {
        const int idx1 = (int)v1.indices[p1];
        if (kf1->has_mappoint[idx1]) continue;
        if (only_stereo) continue;  // bStereo1 = (!mpCamera2 && ...) = false                                  :958-960
        const orbx_kp& kp1 = kf1->kps[idx1];
        const bool right1 = idx1 >= n_left1;
        const uint8_t* d1 = kf1->desc + (size_t)idx1 * 32;
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int p2 = v2.offsets[b]; p2 < v2.offsets[b + 1]; p2++) {
          const int idx2 = (int)v2.indices[p2];
          if (kf2->has_mappoint[idx2]) continue;
          const int dist = orbref_descriptor_distance(d1, kf2->desc + (size_t)idx2 * 32);
          if (dist > TH_LOW || dist > bestDist) continue;
          const orbx_kp& kp2 = kf2->kps[idx2];
          const bool right2 = idx2 >= n_left2;
          bool ok = coarse != 0;
          if (!ok) {
            const float* F12 = F12x4 + 9 * (2 * (int)right1 + (int)right2);
            const float ea = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
            const float eb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
            const float ec = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
            const float num = ea * kp2.x + eb * kp2.y + ec;
            const float den = ea * ea + eb * eb;
            if (den != 0) {
              const float dsqr = num * num / den;
              ok = (double)dsqr < 3.84 * (double)kf2->level_sigma2[kp2.octave];
            }
          }
          if (ok) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
          matches12[idx1] = bestIdx2;
          nmatches++;
          if (check_orientation) rotHist[rot_bin(kp1.angle, kf2->kps[bestIdx2].angle)].push_back(idx1);
        }
      }
      a++;
      b++;
    } else if (v1.node_ids[a] < v2.node_ids[b]) {
      while (a < v1.n_nodes && v1.node_ids[a] < v2.node_ids[b]) a++;
    } else {
      while (b < v2.n_nodes && v2.node_ids[b] < v1.node_ids[a]) b++;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, kHisto, ind1, ind2, ind3);
    for (int i = 0; i < kHisto; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) {
        matches12[idx1] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

// The descriptor part of SearchForTriangulation for rigs whose epipolar test the device cannot evaluate (KannalaBrandt8:
// TriangulateMatches): for every kf1 feature without a MapPoint, the kf2 features without a MapPoint under the same
// vocabulary node whose distance is <= TH_LOW, in the reference's scan order (:973-988). offsets[kf1->n + 1]; returns the
// total (writes at most cap entries).
int orbref_triangulation_candidates(const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, int32_t* offsets,
                                    int32_t* cand_idx2, int32_t* cand_dist, int cap) {
  const int TH_LOW = 50;
  std::vector<std::vector<std::pair<int, int>>> rows(kf1->n);
  const orbx_featvec& v1 = kf1->featvec;
  const orbx_featvec& v2 = kf2->featvec;
  int a = 0, b = 0;
  while (a < v1.n_nodes && b < v2.n_nodes) {
    if (v1.node_ids[a] == v2.node_ids[b]) {
      for (int p1 = v1.offsets[a]; p1 < v1.offsets[a + 1]; p1++) {
        const int idx1 = (int)v1.indices[p1];
        if (kf1->has_mappoint[idx1]) continue;
        const uint8_t* d1 = kf1->desc + (size_t)idx1 * 32;
        for (int p2 = v2.offsets[b]; p2 < v2.offsets[b + 1]; p2++) {
          const int idx2 = (int)v2.indices[p2];
          if (kf2->has_mappoint[idx2]) continue;
          const int dist = orbref_descriptor_distance(d1, kf2->desc + (size_t)idx2 * 32);
          if (dist <= TH_LOW) rows[idx1].push_back({idx2, dist});
        }
      }
      a++;
      b++;
    } else if (v1.node_ids[a] < v2.node_ids[b]) {
      while (a < v1.n_nodes && v1.node_ids[a] < v2.node_ids[b]) a++;
    } else {
      while (b < v2.n_nodes && v2.node_ids[b] < v1.node_ids[a]) b++;
    }
  }
  int total = 0;
  for (int i = 0; i < kf1->n; i++) {
    offsets[i] = total;
    for (const auto& c : rows[i]) {
      if (total < cap) {
        cand_idx2[total] = c.first;
        cand_dist[total] = c.second;
      }
      total++;
    }
  }
  offsets[kf1->n] = total;
  return total;
}

}  // extern "C"
