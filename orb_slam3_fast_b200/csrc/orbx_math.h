// orbx_math.h — scalar arithmetic shared by the CUDA kernels and the host-side planner.
//
// Everything here is __host__ __device__ so that tests/hostcheck.cpp can run the SAME source on the CPU against the
// oracle (tests only; the product never computes results on the host). Each function names the reference call site
// whose arithmetic it has to reproduce bit-for-bit (paths relative to the reference checkout).
//
// Float discipline: the reference is built without -march=native (CMakeLists.txt:13-18) => no FMA contraction.
// Device code is compiled with -fmad=false and the parity-critical expressions additionally use the explicit
// round-to-nearest intrinsics, host code with -ffp-contract=off.
#ifndef ORBX_MATH_H_
#define ORBX_MATH_H_

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define ORBX_HD __host__ __device__ __forceinline__
#else
#define ORBX_HD inline
#endif

namespace orbx {

constexpr int kPatchSize = 31;      // PATCH_SIZE       src/ORBextractor.cc:71
constexpr int kHalfPatch = 15;      // HALF_PATCH_SIZE  :72
constexpr int kEdge = 19;           // EDGE_THRESHOLD   :73
constexpr int kMinBorder = 16;      // EDGE_THRESHOLD - 3   :768-769
constexpr float kCellW = 35.f;      // W                :778
constexpr int kMaxLevels = 16;

// ---- non-fused float primitives -------------------------------------------------------------------------------
ORBX_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
ORBX_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
ORBX_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
ORBX_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
ORBX_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
ORBX_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
ORBX_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  volatile double r = a - b;
  return r;
#endif
}

// cvRound(float): SSE cvtss2si under the default rounding mode = round half to even.
ORBX_HD int cv_round(float v) {
#if defined(__CUDA_ARCH__)
  return __float2int_rn(v);
#else
  return (int)lrintf(v);
#endif
}

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
ORBX_HD int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

// cv::fastAtan2(y, x) in degrees — OpenCV core/mathfuncs_core atan_f32; called at src/ORBextractor.cc:98.
ORBX_HD float fast_atan2_deg(float y, float x) {
  // p_i = coefficient(float) * (float)(180/pi), the product rounded to float once (as OpenCV's static consts are)
  const float p1 = 0x1.ca44dep+5f;    //  0.9997878412794807f * 57.29578f
  const float p3 = -0x1.2aaddcp+4f;   // -0.3258083974640975f * 57.29578f
  const float p5 = 0x1.1d3f7ep+3f;    //  0.1555786518463281f * 57.29578f
  const float p7 = -0x1.4515b2p+1f;   // -0.04432655554792128f * 57.29578f
  const float eps = 0x1p-52f;         // (float)DBL_EPSILON
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = fdiv(ay, fadd(ax, eps));
    c2 = fmul(c, c);
    a = fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = fdiv(ax, fadd(ay, eps));
    c2 = fmul(c, c);
    a = fsub(90.f, fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = fsub(180.f, a);
  if (y < 0) a = fsub(360.f, a);
  return a;
}

// glibc 2.39 cosf / sinf (sysdeps/ieee754/flt-32/s_sincosf.h, the ARM optimized-routines algorithm) restated in plain
// double without contraction, for |y| < 120 (the reference only passes angle * pi/180 with angle in [0, 360)).
// Called at src/ORBextractor.cc:107. Returns cos in *c, sin in *s.
ORBX_HD double sincosf_poly(double x, double x2, int n, bool neg_cos_coeffs) {
  // sign of the cosine coefficients flips when n & 2 (table __sincosf_table[1] in glibc)
  const double sgn = neg_cos_coeffs ? -1.0 : 1.0;
  const double C0 = sgn * 0x1p0, C1 = sgn * -0x1.ffffffd0c621cp-2, C2 = sgn * 0x1.55553e1068f19p-5,
               C3 = sgn * -0x1.6c087e89a359dp-10, C4 = sgn * 0x1.99343027bf8c3p-16;
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    const double x3 = dmul(x, x2);
    const double s1 = dadd(S2, dmul(x2, S3));
    const double x7 = dmul(x3, x2);
    const double s = dadd(x, dmul(x3, S1));
    return dadd(s, dmul(x7, s1));
  } else {
    const double x4 = dmul(x2, x2);
    const double c2 = dadd(C3, dmul(x2, C4));
    const double c1 = dadd(C0, dmul(x2, C1));
    const double x6 = dmul(x4, x2);
    const double c = dadd(c1, dmul(x4, C2));
    return dadd(c, dmul(x6, c2));
  }
}

ORBX_HD uint32_t abstop12(float x) {
  union { float f; uint32_t u; } v;
  v.f = x;
  return (v.u >> 20) & 0x7ff;
}

ORBX_HD void sincosf_glibc(float y, float* c_out, float* s_out) {
  double x = (double)y;
  const uint32_t top = abstop12(y);
  if (top < 0x3f4) {  // |y| < pi/4  (abstop12(pio4f) = 0x3f4)
    const double x2 = dmul(x, x);
    if (top < 0x398) {  // |y| < 2^-12
      *c_out = 1.0f;
      *s_out = y;
      return;
    }
    *s_out = (float)sincosf_poly(x, x2, 0, false);
    *c_out = (float)sincosf_poly(x, x2, 1, false);
    return;
  }
  // reduce_fast: |y| < 120
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  const double r = dmul(x, hpi_inv);
  const int n = ((int32_t)r + 0x800000) >> 24;
  x = dsub(x, dmul((double)n, hpi));
  // sinf and cosf share the quadrant sign and the (n & 2) table flip; cosf evaluates the other polynomial (n ^ 1).
  const double sign_tab[4] = {1.0, -1.0, -1.0, 1.0};
  const double sg = sign_tab[n & 3];
  const bool flip = (n & 2) != 0;
  const double xs = dmul(x, sg);
  const double x2 = dmul(x, x);
  *s_out = (float)sincosf_poly(xs, x2, n, flip);
  *c_out = (float)sincosf_poly(xs, x2, n ^ 1, flip);
}

// ---- glibc logf (sysdeps/ieee754/flt-32/e_logf.c, the ARM optimized-routines version glibc ships since 2.27) ------
// Called by MapPoint::PredictScale as log(ratio) on a float (src/MapPoint.cc:566; <cmath>'s float overload) and by the
// Frame constructors for mfLogScaleFactor (src/Frame.cc:189). 16-entry table, order-3 polynomial in double, one final
// rounding. tests/test_host_math.py compares this with the host's logf on every positive finite float pattern.
ORBX_HD float logf_glibc(float x) {
  const double T[16][2] = {
      {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
      {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
      {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
      {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
      {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
      {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
      {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
      {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
  const double Ln2 = 0x1.62e42fefa39efp-1;
  const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
  union { float f; uint32_t u; } v;
  v.f = x;
  uint32_t ix = v.u;
  if (ix == 0x3f800000u) return 0.f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return -INFINITY;
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
    v.f = fmul(x, 0x1p23f);  // subnormal: normalise
    ix = v.u - (23u << 23);
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) % 16);
  const int k = (int32_t)tmp >> 23;
  v.u = ix - (tmp & 0xff800000u);
  const double z = (double)v.f;
  const double r = dsub(dmul(z, T[i][0]), 1.0);
  const double y0 = dadd(T[i][1], dmul((double)k, Ln2));
  const double r2 = dmul(r, r);
  double y = dadd(dmul(A1, r), A2);
  y = dadd(dmul(A0, r2), y);
  y = dadd(dmul(y, r2), dadd(y0, r));
  return (float)y;
}

// ---- Frame::isInFrustum (src/Frame.cc:632-699, Nleft == -1) for one MapPoint -------------------------------------
// Eigen evaluates a fixed-size 3-vector reduction (dot, squaredNorm, a row of Matrix3f * Vector3f) without packets
// (3 floats do not fill a Packet4f) through redux_novec_unroller, which splits the range in halves:
// c0 + (c1 + c2) (Eigen/src/Core/Redux.h); the reference is built without FMA (CMakeLists.txt:13-18).
ORBX_HD float eigen_dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  return fadd(fmul(a0, b0), fadd(fmul(a1, b1), fmul(a2, b2)));
}
ORBX_HD float sqrtf_rn(float v) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(v);
#else
  return sqrtf(v);
#endif
}
struct FrustumOut {
  bool in_view;          // the return value = mbTrackInView
  bool proj_valid;       // mTrackProjX / Y were set to uv (the bounds test passed); otherwise they are -1
  float proj_x, proj_y, proj_xr, depth, view_cos;
  int level;
};
// fr points at the 26 words of an orbx_frustum (include/orbx_types.h); pos / normal at 3 floats each.
ORBX_HD FrustumOut is_in_frustum(const float* fr, int n_levels, const float* pos, const float* normal, float min_dist_raw,
                                 float max_dist_raw, float viewing_cos_limit) {
  FrustumOut o;
  o.in_view = false;
  o.proj_valid = false;
  o.proj_x = o.proj_y = -1.f;
  o.proj_xr = o.depth = o.view_cos = 0.f;
  o.level = 0;
  const float *R = fr, *t = fr + 9, *Ow = fr + 12;
  const float fx = fr[15], fy = fr[16], cx = fr[17], cy = fr[18], mbf = fr[19];
  const float min_x = fr[20], max_x = fr[21], min_y = fr[22], max_y = fr[23], log_sf = fr[24];
  // Pc = mRcw * P + mtcw (:642)
  const float pcx = fadd(eigen_dot3(R[0], R[1], R[2], pos[0], pos[1], pos[2]), t[0]);
  const float pcy = fadd(eigen_dot3(R[3], R[4], R[5], pos[0], pos[1], pos[2]), t[1]);
  const float pcz = fadd(eigen_dot3(R[6], R[7], R[8], pos[0], pos[1], pos[2]), t[2]);
  const float pc_dist = sqrtf_rn(eigen_dot3(pcx, pcy, pcz, pcx, pcy, pcz));  // Pc.norm() (:643)
  const float invz = fdiv(1.0f, pcz);                                         // :647
  if (pcz < 0.0f) return o;                                                   // :648
  const float u = fadd(fdiv(fmul(fx, pcx), pcz), cx);                         // Pinhole::project, Pinhole.cpp:47-53
  const float v = fadd(fdiv(fmul(fy, pcy), pcz), cy);
  if (u < min_x || u > max_x) return o;                                       // :652-653
  if (v < min_y || v > max_y) return o;
  o.proj_valid = true;
  o.proj_x = u;
  o.proj_y = v;
  const float maxDistance = fmul(1.2f, max_dist_raw), minDistance = fmul(0.8f, min_dist_raw);  // MapPoint.cc:533-541
  const float pox = fsub(pos[0], Ow[0]), poy = fsub(pos[1], Ow[1]), poz = fsub(pos[2], Ow[2]);
  const float dist = sqrtf_rn(eigen_dot3(pox, poy, poz, pox, poy, poz));      // :662
  if (dist < minDistance || dist > maxDistance) return o;                    // :664
  const float viewCos = fdiv(eigen_dot3(pox, poy, poz, normal[0], normal[1], normal[2]), dist);  // :669
  if (viewCos < viewing_cos_limit) return o;                                  // :671
  // MapPoint::PredictScale(dist, this) (src/MapPoint.cc:559-573): ceil(logf(mfMaxDistance / dist) / mfLogScaleFactor)
  const float ratio = fdiv(max_dist_raw, dist);
  int nScale = (int)ceilf(fdiv(logf_glibc(ratio), log_sf));
  if (nScale < 0) nScale = 0;
  else if (nScale >= n_levels) nScale = n_levels - 1;
  o.in_view = true;
  o.proj_xr = fsub(u, fmul(mbf, invz));                                       // :680
  o.depth = pc_dist;
  o.level = nScale;
  o.view_cos = viewCos;
  return o;
}

// ---- libstdc++ std::sort (introsort + final insertion sort, _S_threshold = 16) ------------------------------------
// The reference sorts vector<pair<int, ExtractorNode*>> with compareNodes (src/ORBextractor.cc:542-555, :686-688).
// compareNodes orders by (size, UL.x); elements equal under it are permuted by the algorithm, and that permutation
// decides which node is split last (SURVEY.md hard part 1), so the algorithm of bits/stl_algo.h / stl_heap.h is
// emulated step by step. An element is {key, id}: key = size << 12 | UL.x (UL.x < 4096), id = node slot.
struct SortElem {
  uint32_t key;
  uint32_t id;
};

ORBX_HD bool se_less(const SortElem& a, const SortElem& b) { return a.key < b.key; }
ORBX_HD void se_swap(SortElem& a, SortElem& b) {
  SortElem t = a;
  a = b;
  b = t;
}

ORBX_HD void ss_push_heap(SortElem* first, int hole, int top, SortElem value) {
  int parent = (hole - 1) / 2;
  while (hole > top && se_less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

ORBX_HD void ss_adjust_heap(SortElem* first, int hole, int len, SortElem value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (se_less(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  ss_push_heap(first, hole, top, value);
}

ORBX_HD void ss_make_heap(SortElem* first, int len) {
  if (len < 2) return;
  int parent = (len - 2) / 2;
  for (;;) {
    SortElem v = first[parent];
    ss_adjust_heap(first, parent, len, v);
    if (parent == 0) return;
    parent--;
  }
}

// std::__partial_sort(first, last, last): heap_select degenerates to make_heap, then sort_heap
ORBX_HD void ss_heap_sort(SortElem* first, int len) {
  ss_make_heap(first, len);
  int last = len;
  while (last > 1) {
    --last;
    // __pop_heap(first, last, last): value = *result; *result = *first; adjust_heap(first, 0, last - first, value)
    SortElem v = first[last];
    first[last] = first[0];
    ss_adjust_heap(first, 0, last, v);
  }
}

ORBX_HD void ss_unguarded_linear_insert(SortElem* a, int last) {
  SortElem val = a[last];
  int next = last - 1;
  while (se_less(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}

ORBX_HD void ss_insertion_sort(SortElem* a, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (se_less(a[i], a[first])) {
      SortElem val = a[i];
      for (int k = i; k > first; --k) a[k] = a[k - 1];
      a[first] = val;
    } else {
      ss_unguarded_linear_insert(a, i);
    }
  }
}

ORBX_HD void ss_move_median_to_first(SortElem* a, int result, int ia, int ib, int ic) {
  if (se_less(a[ia], a[ib])) {
    if (se_less(a[ib], a[ic])) se_swap(a[result], a[ib]);
    else if (se_less(a[ia], a[ic])) se_swap(a[result], a[ic]);
    else se_swap(a[result], a[ia]);
  } else if (se_less(a[ia], a[ic])) se_swap(a[result], a[ia]);
  else if (se_less(a[ib], a[ic])) se_swap(a[result], a[ic]);
  else se_swap(a[result], a[ib]);
}

ORBX_HD int ss_unguarded_partition(SortElem* a, int first, int last, int pivot) {
  for (;;) {
    while (se_less(a[first], a[pivot])) ++first;
    --last;
    while (se_less(a[pivot], a[last])) --last;
    if (!(first < last)) return first;
    se_swap(a[first], a[last]);
    ++first;
  }
}

// std::sort(a, a + n, compareNodes). `stack` needs 3 * (2*floor(log2 n) + 2) ints (kSortStack covers n < 2^24).
#ifndef ORBX_SORT_HEAP_HOOK
#define ORBX_SORT_HEAP_HOOK (void)0  // tests count how often the depth limit fires
#endif
constexpr int kSortStack = 160;
ORBX_HD void std_sort_emulate(SortElem* a, int n, int* stack) {
  if (n <= 0) return;
  int lg = 0;
  for (int t = n; t > 1; t >>= 1) lg++;
  // __introsort_loop with the recursion on [cut, last) turned into an explicit stack (recursion happens first,
  // i.e. depth-first into the right part, exactly as the recursive original).
  int sp = 0;
  stack[sp++] = 0;
  stack[sp++] = n;
  stack[sp++] = 2 * lg;
  while (sp > 0) {
    int depth = stack[--sp];
    int last = stack[--sp];
    int first = stack[--sp];
    while (last - first > 16) {
      if (depth == 0) {
        ORBX_SORT_HEAP_HOOK;
        ss_heap_sort(a + first, last - first);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      ss_move_median_to_first(a, first, first + 1, mid, last - 1);
      const int cut = ss_unguarded_partition(a, first + 1, last, first);
      // original: recurse on [cut, last) NOW, then continue the loop on [first, cut).
      // Equivalent order of effects: ranges are disjoint, so processing [first, cut) later is identical.
      stack[sp++] = first;
      stack[sp++] = cut;
      stack[sp++] = depth;
      first = cut;
    }
  }
  // __final_insertion_sort
  if (n > 16) {
    ss_insertion_sort(a, 0, 16);
    for (int i = 16; i < n; ++i) ss_unguarded_linear_insert(a, i);
  } else {
    ss_insertion_sort(a, 0, n);
  }
}

}  // namespace orbx

#endif  // ORBX_MATH_H_
