// k_quadtree.cu — launches orbx::quadtree_run (orbx_quadtree.h): one CTA of kQtWarps warps per (frame, level); warp 0
// runs the tree, the others help with everything that is per candidate (the gather below, the sweeps of the tree).
// Before the tree: the per-cell candidate slots written by k_fast are compacted, cells in row-major order, into the
// level's candidate array — the order ComputeKeyPointsOctTree appends them in (src/ORBextractor.cc:905-958). The
// array (4 B per candidate) and the per-candidate node labels (2 B) live in shared memory when the level has at most
// kSmemCand candidates (the normal case: ~1.5 k at level 0 of a textured 752x480 frame); noise-like images fall back
// to the global scratch — same code, different pointers.
// After the tree: the selected candidates are stored in list order, and — the first half of
// ORBextractor::operator()'s output assembly (:1083-1101) — every keypoint gets its rank among the level's "mono" or
// "stereo" (lapping area) keypoints, so that k_describe can place it without a separate assembly pass.
#include "orbx_kernels.cuh"
#include "orbx_quadtree.h"

namespace orbx {

constexpr int kSmemCand = 2048;
constexpr int kQtWarps = 4;  // tools/qt_prof.py: 70 % of a level-0 tree (125 us with one warp) is per-candidate work

struct QtSmem {
  int cap;
  size_t off_box0, off_box1, off_cnt0, off_cnt1, off_ch0, off_ch1, off_newpos, off_childpos, off_committed,
      off_splittable, off_pend0, off_pend1, off_sort, off_rank, off_scan, off_vars, off_cand, off_lab, total;
};

static QtSmem qt_layout(const Plan& P) {
  QtSmem s;
  int cap = 8;
  for (int l = 0; l < P.nlevels; l++)
    if (P.lv[l].kp_cap > cap) cap = P.lv[l].kp_cap;
  s.cap = cap;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 15) / 16 * 16;
    return r;
  };
  s.off_box0 = take(sizeof(QBox) * cap);
  s.off_box1 = take(sizeof(QBox) * cap);
  s.off_cnt0 = take(4 * cap);
  s.off_cnt1 = take(4 * cap);
  s.off_ch0 = take(16 * cap);
  s.off_ch1 = take(16 * cap);
  s.off_newpos = take(2 * cap);
  s.off_childpos = take(8 * cap);
  s.off_committed = take(cap);
  s.off_splittable = take(2 * cap);
  s.off_pend0 = take(2 * cap);
  s.off_pend1 = take(2 * cap);
  s.off_sort = take(sizeof(SortElem) * cap);
  s.off_rank = take(2 * cap);
  s.off_scan = take(4 * (cap + 1));
  s.off_vars = take(32);
  s.off_cand = take(4 * kSmemCand);
  s.off_lab = take(2 * kSmemCand);
  s.total = o;
  return s;
}

size_t quadtree_smem_bytes(const Plan& P) { return qt_layout(P).total; }

__global__ void __launch_bounds__(kQtWarps * 32) k_quadtree(const __grid_constant__ Plan P, const WorkSet ws, const QtSmem S,
                                                 int lap0, int lap1) {
  extern __shared__ __align__(16) uint8_t smem[];
  // blockIdx.y = level: blocks are handed out level 0 first, so the long trees (level 0: ~100 us, level 7: ~8 us) start
  // first and the tail of the launch is made of short ones (longest-processing-time-first)
  const int l = blockIdx.y, f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const LevelPlan& L = P.lv[l];
  const uint32_t* slots = ws.slots + (int64_t)f * P.slots_per_frame + L.slot_base;
  const int32_t* counts = ws.cell_count + (int64_t)f * P.cells_per_frame + L.cell_base;

  QTree T;
  T.prof = nullptr;
  T.prof_prev = 0;
#ifdef ORBX_QT_PROF
  if (ws.qt_prof && warp == 0) {
    T.prof = ws.qt_prof + ((int64_t)f * P.nlevels + l) * kQtProfSlots;
    if (lane == 0)
      for (int k = 0; k < kQtProfSlots; k++) T.prof[k] = 0;
    T.prof_prev = clock64();
  }
#endif
  // ---- how many candidates? (decides where they live) ----
  const int ncell = L.nCols * L.nRows;
  int C = 0;
  for (int i = lane; i < ncell; i += 32) C += counts[i];
  C = __reduce_add_sync(0xffffffffu, C);
  uint32_t* cand = reinterpret_cast<uint32_t*>(smem + S.off_cand);
  uint16_t* lab = reinterpret_cast<uint16_t*>(smem + S.off_lab);
  if (C > kSmemCand) {
    cand = ws.cand + (int64_t)f * P.slots_per_frame + L.slot_base;
    lab = ws.lab + (int64_t)f * P.slots_per_frame + L.slot_base;
  }

  // ---- compact the cell slots in cell order: exclusive scan of the cell counts (in chunks), then a flat gather —
  //      output o belongs to the last cell whose prefix is <= o (binary search in shared memory). Every lane step is
  //      independent, 4 are in flight per lane. The prefix array borrows the tree's node arrays (not live yet). ----
  int* pref = reinterpret_cast<int*>(smem + S.off_box0);
  int* vars = reinterpret_cast<int*>(smem + S.off_vars);
  const int kCellChunk = (int)((S.off_newpos - S.off_box0) / 4) - 1;
  int done = 0;
  for (int cb = 0; cb < ncell; cb += kCellChunk) {
    const int nc = min(kCellChunk, ncell - cb);
    if (warp == 0) {
      for (int i = lane; i < nc; i += 32) pref[i] = counts[cb + i];
      __syncwarp();
      const int t = excl_scan(pref, nc);
      if (lane == 0) vars[kQtVarTotal] = t;
    }
    __syncthreads();
    const int total = vars[kQtVarTotal];
    auto locate = [&](int o) {
      int lo = 0, hi = nc - 1;  // largest i with pref[i] <= o
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (pref[mid] <= o) lo = mid;
        else hi = mid - 1;
      }
      return slots + (int64_t)(cb + lo) * L.slot_cap + (o - pref[lo]);
    };
    const int tn = kQtWarps * 32;
    for (int o0 = threadIdx.x; o0 < total; o0 += 4 * tn) {
      uint32_t v[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (o0 + tn * u < total) v[u] = *locate(o0 + tn * u);
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (o0 + tn * u < total) cand[done + o0 + tn * u] = v[u];
    }
    done += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) ws.lvl_c[f * P.nlevels + l] = C;
  ORBX_QT_MARK(T, 1);

  T.cap = S.cap;
  T.box[0] = reinterpret_cast<QBox*>(smem + S.off_box0);
  T.box[1] = reinterpret_cast<QBox*>(smem + S.off_box1);
  T.cnt[0] = reinterpret_cast<int*>(smem + S.off_cnt0);
  T.cnt[1] = reinterpret_cast<int*>(smem + S.off_cnt1);
  T.child[0] = reinterpret_cast<int*>(smem + S.off_ch0);
  T.child[1] = reinterpret_cast<int*>(smem + S.off_ch1);
  T.newpos = reinterpret_cast<uint16_t*>(smem + S.off_newpos);
  T.childpos = reinterpret_cast<uint16_t*>(smem + S.off_childpos);
  T.committed = smem + S.off_committed;
  T.splittable = smem + S.off_splittable;
  T.pending[0] = reinterpret_cast<uint16_t*>(smem + S.off_pend0);
  T.pending[1] = reinterpret_cast<uint16_t*>(smem + S.off_pend1);
  T.sortbuf = reinterpret_cast<SortElem*>(smem + S.off_sort);
  T.rank2pos = reinterpret_cast<uint16_t*>(smem + S.off_rank);
  T.scan = reinterpret_cast<int*>(smem + S.off_scan);
  T.vars = reinterpret_cast<int*>(smem + S.off_vars);
  T.cand = cand;
  T.lab = lab;
  T.C = C;
  uint32_t* out = reinterpret_cast<uint32_t*>(smem + S.off_sort);  // the sort buffer is free once the tree is built

  const int width = L.maxBX - kMinBorder, height = L.maxBY - kMinBorder;
  if (warp != 0) {  // helper warps: their share of the candidate sweeps, until warp 0 releases them
    quadtree_helper(T, L.hX);
    return;
  }
  const int n = quadtree_run(T, width, height, L.nIni, L.hX, L.quota, out);
  quadtree_release(T);

  // ---- selected keypoints in list order + their rank among the level's mono / stereo keypoints ----
  uint32_t* kp = ws.lvl_kp + (int64_t)f * P.kps_per_frame + L.kp_base;
  int32_t* dst = ws.dst + (int64_t)f * P.kps_per_frame + L.kp_base;
  const float flap0 = (float)lap0, flap1 = (float)lap1;
  const unsigned lt = (1u << lane) - 1u;
  int n_st = 0, n_mo = 0;
  for (int base = 0; base < n; base += 32) {
    const int p = base + lane;
    bool is_st = false;
    const bool valid = p < n;
    uint32_t cw = 0;
    if (valid) {
      cw = cand[out[p]];
      float x = (float)(cand_x(cw) + kMinBorder);
      if (l != 0) x = fmul(x, L.scale);           // keypoint->pt *= scale          :1086
      is_st = x >= flap0 && x <= flap1;           // inclusive lapping test          :1088-1089
    }
    const unsigned m_st = __ballot_sync(0xffffffffu, valid && is_st);
    const unsigned m_mo = __ballot_sync(0xffffffffu, valid && !is_st);
    if (valid) {
      kp[p] = cw;
      dst[p] = is_st ? (int32_t)(0x40000000u | (uint32_t)(n_st + __popc(m_st & lt))) : n_mo + __popc(m_mo & lt);
    }
    n_st += __popc(m_st);
    n_mo += __popc(m_mo);
  }
  if (lane == 0) {
    ws.lvl_n[f * P.nlevels + l] = n;
    ws.lvl_st[f * P.nlevels + l] = n_st;
  }
  ORBX_QT_MARK(T, 9);
#ifdef ORBX_QT_PROF
  if (lane == 0 && T.prof) {
    T.prof[12] = C;
    T.prof[13] = n;
  }
#endif
}

void launch_quadtree(const Plan& P, const WorkSet& ws, int lap0, int lap1, int frames, cudaStream_t st) {
  const QtSmem S = qt_layout(P);
  cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(S.total > 48 * 1024 ? S.total : 48 * 1024));
  dim3 grid(frames, P.nlevels);
  k_quadtree<<<grid, kQtWarps * 32, S.total, st>>>(P, ws, S, lap0, lap1);
}

}  // namespace orbx
