// k_quadtree.cu — launches orbx::quadtree_run (orbx_quadtree.h) with one warp per (frame, level).
// Before the tree: the per-cell candidate slots written by k_fast are compacted, cells in row-major order, into the
// level's candidate array — the order ComputeKeyPointsOctTree appends them in (src/ORBextractor.cc:905-958).
// After the tree: the selected candidates are stored in list order for the orientation / descriptor stage.
#include "orbx_kernels.cuh"
#include "orbx_quadtree.h"

namespace orbx {

struct QtSmem {
  int cap;
  size_t off_box0, off_box1, off_cnt0, off_cnt1, off_ch0, off_ch1, off_newpos, off_childpos, off_committed,
      off_splittable, off_pend0, off_pend1, off_sort, off_rank, off_scan, off_vars, off_out, total;
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
  s.off_out = take(4 * cap);
  s.total = o;
  return s;
}

size_t quadtree_smem_bytes(const Plan& P) { return qt_layout(P).total; }

__global__ void __launch_bounds__(32) k_quadtree(const __grid_constant__ Plan P, const WorkSet ws, const QtSmem S) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int l = blockIdx.x, f = blockIdx.y, lane = threadIdx.x;
  const LevelPlan& L = P.lv[l];
  const uint32_t* slots = ws.slots + (int64_t)f * P.slots_per_frame + L.slot_base;
  const int32_t* counts = ws.cell_count + (int64_t)f * P.cells_per_frame + L.cell_base;
  uint32_t* cand = ws.cand + (int64_t)f * P.slots_per_frame + L.slot_base;
  uint32_t* lab = ws.lab + (int64_t)f * P.slots_per_frame + L.slot_base;

  // ---- compact the cell slots in cell order ----
  const int ncell = L.nCols * L.nRows;
  int C = 0;
  for (int base = 0; base < ncell; base += 32) {
    const int ci = base + lane;
    const int n = ci < ncell ? counts[ci] : 0;
    int inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    const int off = C + inc - n;
    const int lim = min(32, ncell - base);
    for (int t = 0; t < lim; t++) {
      const int nt = __shfl_sync(0xffffffffu, n, t);
      const int ot = __shfl_sync(0xffffffffu, off, t);
      const uint32_t* s = slots + (int64_t)(base + t) * L.slot_cap;
      for (int k = lane; k < nt; k += 32) cand[ot + k] = s[k];
    }
    C += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
  if (lane == 0) ws.lvl_c[f * P.nlevels + l] = C;

  QTree T;
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
  uint32_t* out = reinterpret_cast<uint32_t*>(smem + S.off_out);

  const int width = L.maxBX - kMinBorder, height = L.maxBY - kMinBorder;
  const int n = quadtree_run(T, width, height, L.nIni, L.hX, L.quota, out);

  uint32_t* kp = ws.lvl_kp + (int64_t)f * P.kps_per_frame + L.kp_base;
  for (int p = lane; p < n; p += 32) kp[p] = cand[out[p]];
  if (lane == 0) ws.lvl_n[f * P.nlevels + l] = n;
}

void launch_quadtree(const Plan& P, const WorkSet& ws, int frames, cudaStream_t st) {
  const QtSmem S = qt_layout(P);
  cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(S.total > 48 * 1024 ? S.total : 48 * 1024));
  dim3 grid(P.nlevels, frames);
  k_quadtree<<<grid, 32, S.total, st>>>(P, ws, S);
}

}  // namespace orbx
