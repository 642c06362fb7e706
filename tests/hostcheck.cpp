// hostcheck.cpp — TEST INFRASTRUCTURE. Compiles the product's __host__ __device__ arithmetic header
// (orb_slam3_fast_b200/csrc/orbx_math.h) for the CPU so that `-m "not gpu"` tests can compare the exact source the
// kernels use against the oracle / glibc / libstdc++ without a GPU. Not part of the product path.
#include <cstdint>
#include <cstring>
#include <vector>

static long g_heap_calls = 0;
#define ORBX_SORT_HEAP_HOOK (++g_heap_calls)
#include "../orb_slam3_fast_b200/csrc/orbx_math.h"
#include "../orb_slam3_fast_b200/csrc/orbx_plan.h"
#include "../orb_slam3_fast_b200/csrc/orbx_quadtree.h"

extern "C" {
float hc_fast_atan2(float y, float x) { return orbx::fast_atan2_deg(y, x); }
void hc_sincosf(float a, float* c, float* s) { orbx::sincosf_glibc(a, c, s); }
long hc_heap_calls() { return g_heap_calls; }
int hc_cv_round(float v) { return orbx::cv_round(v); }
// std::sort emulation: keys (size, ulx) -> permutation
void hc_std_sort_perm(const int* key0, const int* key1, int n, int* perm) {
  std::vector<orbx::SortElem> a(n);
  for (int i = 0; i < n; i++) {
    a[i].key = ((uint32_t)key0[i] << 12) | (uint32_t)key1[i];
    a[i].id = (uint32_t)i;
  }
  int stack[orbx::kSortStack];
  orbx::std_sort_emulate(a.data(), n, stack);
  for (int i = 0; i < n; i++) perm[i] = (int)a[i].id;
}
// the warp-cooperative std::sort emulation (one "lane" on the CPU): same interface
void hc_std_sort_perm_warp(const int* key0, const int* key1, int n, int* perm) {
  std::vector<orbx::SortElem> a(n > 0 ? n : 1), tmp(n > 0 ? n : 1);
  std::vector<uint16_t> li(n > 0 ? n : 1), ri(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) {
    a[i].key = ((uint32_t)key0[i] << 12) | (uint32_t)key1[i];
    a[i].id = (uint32_t)i;
  }
  orbx::SortScratch W{li.data(), ri.data(), tmp.data()};
  orbx::std_sort_emulate_warp(a.data(), n, W);
  for (int i = 0; i < n; i++) perm[i] = (int)a[i].id;
}
// counts mismatches of sincosf_glibc vs glibc over all floats with bit patterns in [lo, hi)
long hc_sincosf_sweep(uint32_t lo, uint32_t hi) {
  long bad = 0;
  for (uint32_t u = lo; u < hi; u++) {
    float y, c, s;
    memcpy(&y, &u, 4);
    orbx::sincosf_glibc(y, &c, &s);
    float rc = cosf(y), rs = sinf(y);
    if (memcmp(&c, &rc, 4) || memcmp(&s, &rs, 4)) bad++;
  }
  return bad;
}
// counts mismatches of logf_glibc vs the host's logf over all floats with bit patterns in [lo, hi) taking every step-th
long hc_logf_sweep(uint32_t lo, uint32_t hi, uint32_t step) {
  long bad = 0;
  for (uint64_t u = lo; u < hi; u += step) {
    const uint32_t b = (uint32_t)u;
    float x;
    memcpy(&x, &b, 4);
    const float a = orbx::logf_glibc(x), r = logf(x);
    if (memcmp(&a, &r, 4) && !(a != a && r != r)) bad++;
  }
  return bad;
}
// Frame::isInFrustum through the product's arithmetic for m points of one map; outputs as orbref_is_in_frustum
void hc_is_in_frustum(const float* fr26, const float* pos, const float* normal, const float* min_dist,
                      const float* max_dist, const uint8_t* skip, int m, float cos_limit, uint8_t* in_view, float* px,
                      float* py, float* pxr, int32_t* level, float* vcos, float* depth) {
  int32_t n_levels;
  memcpy(&n_levels, fr26 + 25, 4);
  for (int i = 0; i < m; i++) {
    in_view[i] = 0;
    if (skip && skip[i]) continue;
    const orbx::FrustumOut o =
        orbx::is_in_frustum(fr26, n_levels, pos + 3 * i, normal + 3 * i, min_dist[i], max_dist[i], cos_limit);
    in_view[i] = o.in_view;
    px[i] = o.proj_x;
    py[i] = o.proj_y;
    if (o.in_view) {
      pxr[i] = o.proj_xr;
      level[i] = o.level;
      vcos[i] = o.view_cos;
      depth[i] = o.depth;
    }
  }
}
int hc_plan(int w, int h, int nfeatures, float scale, int nlevels, orbx::Plan* out) {
  return orbx::make_plan(w, h, nfeatures, scale, nlevels, out);
}
int hc_plan_size() { return (int)sizeof(orbx::Plan); }
void hc_axis_table(int ssize, int dsize, int clamp, int16_t* ofs, int16_t* c0, int16_t* c1) {
  orbx::axis_table(ssize, dsize, clamp != 0, ofs, c0, c1);
}

// DistributeOctTree through the product's array algorithm, one "lane". cand = packed x|y<<12|score<<24.
int hc_quadtree(const uint32_t* cand, int C, int width, int height, int nIni, float hX, int N, int cap,
                uint32_t* out_idx) {
  using namespace orbx;
  std::vector<QBox> box0(cap), box1(cap);
  std::vector<int> cnt0(cap), cnt1(cap), ch0(cap * 4), ch1(cap * 4), scan(cap + 1), vars(8);
  std::vector<uint16_t> newpos(cap), childpos(cap * 4), pend0(cap), pend1(cap), rank2pos(cap);
  std::vector<uint8_t> committed(cap), splittable(2 * cap);
  std::vector<SortElem> sortbuf(cap);
  std::vector<uint16_t> lab(C > 0 ? C : 1);
  QTree T;
  T.cap = cap;
  T.box[0] = box0.data(); T.box[1] = box1.data();
  T.cnt[0] = cnt0.data(); T.cnt[1] = cnt1.data();
  T.child[0] = ch0.data(); T.child[1] = ch1.data();
  T.newpos = newpos.data(); T.childpos = childpos.data();
  T.committed = committed.data(); T.splittable = splittable.data();
  T.pending[0] = pend0.data(); T.pending[1] = pend1.data();
  T.sortbuf = sortbuf.data(); T.rank2pos = rank2pos.data();
  T.scan = scan.data(); T.vars = vars.data();
  T.cand = cand; T.lab = lab.data(); T.C = C;
  return quadtree_run(T, width, height, nIni, hX, N, out_idx);
}
}
