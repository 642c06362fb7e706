// orbm_api.cu — the extern "C" matcher ABI declared in include/orbm.h: uploads the flat views, launches the kernels
// of k_match.cu / k_search.cu on the context's stream, downloads the results. No result is computed on the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/orbm.h"
#include "orbm_handle.h"
#include "orbx_handle.h"
#include "orbx_match.cuh"

using namespace orbx;

namespace orbm_detail {
std::string& create_error() {
  thread_local std::string e;
  return e;
}
}  // namespace orbm_detail
#define g_m_create_error (orbm_detail::create_error())

namespace {

void pyr_view(const orbx_extractor* ex, PyrView* v, int ln) {
  memset(v, 0, sizeof(*v));
  const Plan& P = ex->plan;
  const FrameSet& fs = ex->lane[ln].last_fs;
  for (int l = 0; l < P.nlevels; l++) {
    if (l == 0) {
      v->base[l] = fs.lvl0;
      v->pitch[l] = fs.pitch0;
      v->fstride[l] = fs.fstride0;
    } else {
      v->base[l] = fs.pyr + P.lv[l].img_off;
      v->pitch[l] = P.lv[l].pitch;
      v->fstride[l] = fs.slab_fstride;
    }
    v->w[l] = P.lv[l].w;
    v->h[l] = P.lv[l].h;
  }
}

// lnl / lnr: the lanes of the two extractors that hold the pair(s)
int stereo_args_common(orbm_matcher* m, const orbx_extractor* left, const orbx_extractor* right, StereoArgs* A,
                       int lnl, int lnr) {
  if (!left || !right || !left->planned || !right->planned || left->lane[lnl].last_frames < 1 || right->lane[lnr].last_frames < 1)
    return mfail(m, ORBX_E_ARG, "stereo match needs both extractors to have run");
  if (left->device != m->device || right->device != m->device)
    return mfail(m, ORBX_E_ARG, "extractors and matcher must live on the same device");
  if (left->nlevels != right->nlevels) return mfail(m, ORBX_E_ARG, "level count mismatch");
  memset(A, 0, sizeof(*A));
  pyr_view(left, &A->left, lnl);
  pyr_view(right, &A->right, lnr);
  A->nlevels = left->nlevels;
  for (int l = 0; l < left->nlevels; l++) {
    A->scale[l] = left->plan.lv[l].scale;
    A->inv_scale[l] = left->plan.lv[l].inv_scale;
  }
  return ORBX_OK;
}

DevFrame upload_frame(Arena& ar, const orbx_frame_view* f) {
  DevFrame F{};
  const int cells = ORBX_GRID_COLS * ORBX_GRID_ROWS;
  F.n = f->n;
  F.n_levels = f->n_levels;
  F.kps = ar.upload(f->kps, f->n);
  F.desc = ar.upload(f->desc, (size_t)f->n * 32);
  F.u_right = f->u_right ? ar.upload(f->u_right, f->n) : (ar.alloc<float>(1), nullptr);
  F.occupied = ar.upload(f->occupied, f->n);
  F.cell_offsets = ar.upload(f->grid.cell_offsets, cells + 1);
  F.cell_items = ar.upload(f->grid.cell_items, (size_t)f->grid.cell_offsets[cells]);
  F.min_x = f->grid.min_x;
  F.min_y = f->grid.min_y;
  F.inv_w = f->grid.inv_w;
  F.inv_h = f->grid.inv_h;
  F.scale_factors = ar.upload(f->scale_factors, f->n_levels);
  return F;
}

// count -> scan -> (size scratch) -> fill -> resolve, then the results come back
int run_search(orbm_matcher* m, Arena& ar, const DevFrame& F, const DevQueries& Q, ResolveArgs R, int n,
               int32_t* assign, int32_t* nmatches, int32_t* decisions = nullptr) {
  cudaStream_t st = m->stream;
  SearchScratch S{};
  S.counts = ar.alloc<int32_t>((size_t)Q.m + 1);
  S.pre = ar.alloc<int4>(Q.m);
  int32_t* d_total = ar.alloc<int32_t>(1);
  R.assign = ar.alloc<int32_t>(n);
  R.nmatches = ar.alloc<int32_t>(1);
  R.events = ar.alloc<int32_t>(2 * (size_t)Q.m);
  R.dec = ar.alloc<int32_t>(Q.m);
  R.occ = nullptr;
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  int32_t total = 0;
  if (Q.m > 0) {
    launch_search_count(F, Q, S.counts, st);
    launch_scan(S.counts, Q.m, d_total, st);
    ORBM_CUDA(m, cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, st));
    ORBM_CUDA(m, cudaStreamSynchronize(st));
  } else {
    ORBM_CUDA(m, cudaMemsetAsync(S.counts, 0, 4, st));
  }
  S.cand_idx = ar.alloc<int32_t>(total);
  S.cand_dist = ar.alloc<int32_t>(total);
  S.cap_total = total;
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_search_fill(F, Q, S, st);
  launch_search_resolve(F, Q, S, R, st);
  ORBM_CUDA(m, cudaGetLastError());
  if (n > 0 && assign) ORBM_CUDA(m, cudaMemcpyAsync(assign, R.assign, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (decisions && Q.m > 0)  // the keypoint every point takes (mode 1), before any orientation filter
    ORBM_CUDA(m, cudaMemcpyAsync(decisions, R.dec, (size_t)Q.m * 4, cudaMemcpyDeviceToHost, st));
  int32_t nm = 0;
  ORBM_CUDA(m, cudaMemcpyAsync(&nm, R.nmatches, 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaStreamSynchronize(st));
  if (nmatches) *nmatches = nm;
  return ORBX_OK;
}

}  // namespace

extern "C" {

int orbm_create(orbm_matcher** out, int device) {
  if (!out) return mfail(nullptr, ORBX_E_ARG, "out == NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return mfail(nullptr, ORBX_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return mfail(nullptr, ORBX_E_ARG, "bad device ordinal");
  orbm_matcher* m = new orbm_matcher;
  m->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_m_create_error = cudaGetErrorString(e);
    delete m;
    return ORBX_E_CUDA;
  }
  *out = m;
  return ORBX_OK;
}

void orbm_destroy(orbm_matcher* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (auto& b : m->buf) b.release();
  for (auto& lb : m->lane_buf)
    for (auto& b : lb) b.release();
  for (auto& h : m->lane_h_nm)
    if (h) cudaFreeHost(h);
  for (auto& b : m->voc_buf) b.release();
  for (auto& b : m->tc_buf) b.release();
  for (auto& lb : m->track_buf)
    for (auto& b : lb) b.release();
  for (auto& b : m->track_map) b.release();
  for (auto& h : m->lane_h_track)
    if (h) cudaFreeHost(h);
  for (auto& sl : m->sgraph)
    if (sl.exec) cudaGraphExecDestroy(sl.exec);
  if (m->sgraph_fork) cudaEventDestroy(m->sgraph_fork);
  if (m->track_map_ready) cudaEventDestroy(m->track_map_ready);
  if (m->up_stream) cudaStreamDestroy(m->up_stream);
  for (cudaEvent_t e : m->up_small)
    if (e) cudaEventDestroy(e);
  if (m->h_stage) cudaFreeHost(m->h_stage);
  if (m->d_stage) cudaFree(m->d_stage);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

const char* orbm_last_error(const orbm_matcher* m) { return m ? m->err.c_str() : g_m_create_error.c_str(); }

int orbm_descriptor_distance_batch(orbm_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist) {
  if (!m || n < 0 || (n > 0 && (!a || !b || !dist))) return mfail(m, ORBX_E_ARG, "bad argument");
  if (n == 0) return ORBX_OK;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  const uint8_t* da = ar.upload(a, (size_t)n * 32);
  const uint8_t* db = ar.upload(b, (size_t)n * 32);
  int32_t* dd = ar.alloc<int32_t>(n);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_desc_dist(da, db, n, dd, m->stream);
  ORBM_CUDA(m, cudaMemcpyAsync(dist, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  return ORBX_OK;
}

int orbm_distinctive_descriptors(orbm_matcher* m, const uint8_t* desc, const int32_t* offsets, int n_points,
                                 int32_t* best_idx) {
  if (!m || n_points < 0 || (n_points > 0 && (!offsets || !best_idx))) return mfail(m, ORBX_E_ARG, "bad argument");
  if (n_points == 0) return ORBX_OK;
  const int total = offsets[n_points];
  if (total < 0 || (total > 0 && !desc)) return mfail(m, ORBX_E_ARG, "bad argument");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  const uint8_t* dd = ar.upload(desc, (size_t)total * 32);
  const int32_t* doff = ar.upload(offsets, (size_t)n_points + 1);
  int32_t* db = ar.alloc<int32_t>(n_points);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_distinctive(dd, doff, n_points, db, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  ORBM_CUDA(m, cudaMemcpyAsync(best_idx, db, (size_t)n_points * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  return ORBX_OK;
}

int orbm_knn2_device(orbm_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int32_t* d_idx1,
                     int32_t* d_d1, int32_t* d_idx2, int32_t* d_d2, void* cuda_stream) {
  if (!m || nq < 0 || nt < 0) return mfail(m, ORBX_E_ARG, "bad argument");
  if (nq == 0) return ORBX_OK;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : m->stream;
  // the partial buffer is the LAST arena slot so that host-API uploads (slots 0..) are not disturbed
  DevBuf& pb = m->buf[kBufs - 1];
  if (knn2_tc_enabled() && knn2_tc_eligible(nq, nt)) {
    // large sets: Hamming as an s8 GEMM on the tensor cores (k_knn2_tc.cu); same keys, same merge, same results
    int tiles_per_split = 0;
    const int splits = knn2_tc_splits(nq, nt, &tiles_per_split);
    cudaError_t e = pb.reserve((size_t)splits * nq * sizeof(int4));
    if (e == cudaSuccess && knn2_tc_expands_queries()) e = m->tc_buf[0].reserve(knn2_tc_expanded_bytes(nq));
    if (e == cudaSuccess) e = m->tc_buf[1].reserve(knn2_tc_expanded_bytes(nt));
    if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
    ORBM_CUDA(m, launch_knn2_tc(d_q, nq, d_t, nt, static_cast<int8_t*>(m->tc_buf[0].p),
                                static_cast<int8_t*>(m->tc_buf[1].p), reinterpret_cast<int4*>(pb.p), splits,
                                tiles_per_split, st));
    launch_knn2_merge(reinterpret_cast<int4*>(pb.p), nq, splits, d_idx1, d_d1, d_idx2, d_d2, st);
    ORBM_CUDA(m, cudaGetLastError());
    return ORBX_OK;
  }
  const int splits = knn2_splits(nq, nt);
  cudaError_t e = pb.reserve((size_t)splits * nq * sizeof(int4));
  if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
  launch_knn2(d_q, nq, d_t, nt, reinterpret_cast<int4*>(pb.p), splits, d_idx1, d_d1, d_idx2, d_d2, st);
  ORBM_CUDA(m, cudaGetLastError());
  return ORBX_OK;
}

int orbm_knn2(orbm_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx1, int32_t* d1,
              int32_t* idx2, int32_t* d2) {
  if (!m || nq < 0 || nt < 0 || (nq > 0 && (!q || !idx1 || !d1 || !idx2 || !d2)) || (nt > 0 && !t))
    return mfail(m, ORBX_E_ARG, "bad argument");
  if (nq == 0) return ORBX_OK;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  const uint8_t* dq = ar.upload(q, (size_t)nq * 32);
  const uint8_t* dt = ar.upload(t, (size_t)nt * 32);
  int32_t* out = ar.alloc<int32_t>((size_t)nq * 4);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  int rc = orbm_knn2_device(m, dq, nq, dt, nt, out, out + nq, out + 2 * (size_t)nq, out + 3 * (size_t)nq, nullptr);
  if (rc) return rc;
  cudaStream_t st = m->stream;
  ORBM_CUDA(m, cudaMemcpyAsync(idx1, out, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(d1, out + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(idx2, out + 2 * (size_t)nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(d2, out + 3 * (size_t)nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaStreamSynchronize(st));
  return ORBX_OK;
}

int orbm_stereo_match_batch_device(orbm_matcher* m, const orbx_extractor* left, const orbx_extractor* right,
                                   int n_pairs, const orbx_kp* d_kps_l, const uint8_t* d_desc_l, const int32_t* d_n_l,
                                   const orbx_kp* d_kps_r, const uint8_t* d_desc_r, const int32_t* d_n_r, int cap,
                                   float mbf, float mb, float* d_u_right, float* d_depth, int32_t* d_n_matched,
                                   void* cuda_stream) {
  if (!m || n_pairs < 1 || cap < 1 || !d_kps_l || !d_desc_l || !d_n_l || !d_kps_r || !d_desc_r || !d_n_r ||
      !d_u_right || !d_depth || !d_n_matched)
    return mfail(m, ORBX_E_ARG, "bad argument");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  StereoArgs A;
  int lnl = 0, lnr = 0, ll = 0, lr = 0;
  if (api_find_frame(left, 0, &lnl, &ll) != ORBX_OK || api_find_frame(right, 0, &lnr, &lr) != ORBX_OK || ll != 0 || lr != 0)
    return mfail(m, ORBX_E_ARG, "stereo match needs both extractors to have run");
  int rc = stereo_args_common(m, left, right, &A, lnl, lnr);
  if (rc) return rc;
  // pairs 0 .. n_pairs - 1 must lie in ONE lane of each extractor (a device-resident call, or a host call of one group)
  if (n_pairs > left->lane[lnl].last_frames || n_pairs > right->lane[lnr].last_frames)
    return mfail(m, ORBX_E_ARG, "n_pairs exceeds the frames of the extractors' last call that are resident in one lane");
  DevBuf& sb = m->buf[kBufs - 2];
  cudaError_t e = sb.reserve((size_t)n_pairs * cap * 4);
  if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
  A.frame0 = 0;
  A.kps_l = d_kps_l; A.kps_r = d_kps_r; A.desc_l = d_desc_l; A.desc_r = d_desc_r;
  A.n_l = d_n_l; A.n_r = d_n_r; A.cap = cap; A.mbf = mbf; A.mb = mb;
  A.u_right = d_u_right; A.depth = d_depth; A.sad = reinterpret_cast<int32_t*>(sb.p); A.n_matched = d_n_matched;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : m->stream;
  launch_stereo(A, n_pairs, cap, st);
  ORBM_CUDA(m, cudaGetLastError());
  return ORBX_OK;
}

int orbm_stereo_match(orbm_matcher* m, const orbx_extractor* left, const orbx_extractor* right, int frame,
                      const orbx_kp* kps_l, const uint8_t* desc_l, int n_l, const orbx_kp* kps_r,
                      const uint8_t* desc_r, int n_r, float mbf, float mb, float* u_right, float* depth,
                      int32_t* n_matched) {
  if (!m || n_l < 0 || n_r < 0 || (n_l > 0 && (!kps_l || !desc_l || !u_right || !depth)) ||
      (n_r > 0 && (!kps_r || !desc_r)))
    return mfail(m, ORBX_E_ARG, "bad argument");
  if (n_matched) *n_matched = 0;
  if (n_l == 0) return ORBX_OK;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  StereoArgs A;
  int lnl = 0, lnr = 0, ll = 0, lr = 0;
  if (api_find_frame(left, frame, &lnl, &ll) != ORBX_OK || api_find_frame(right, frame, &lnr, &lr) != ORBX_OK || ll != lr)
    return mfail(m, ORBX_E_ARG, "frame out of range, or no longer resident (only the last kLanes groups of a call are)");
  int rc = stereo_args_common(m, left, right, &A, lnl, lnr);
  if (rc) return rc;
  // the extractors ran on their own streams
  ORBM_CUDA(m, cudaStreamSynchronize(left->lane[lnl].stream));
  ORBM_CUDA(m, cudaStreamSynchronize(right->lane[lnr].stream));
  Arena ar(m);
  A.frame0 = ll;
  A.kps_l = ar.upload(kps_l, n_l);
  A.desc_l = ar.upload(desc_l, (size_t)n_l * 32);
  A.kps_r = ar.upload(kps_r, n_r);  // n_r == 0: a 1-element dummy is allocated, the count gates every read
  A.desc_r = ar.upload(desc_r, (size_t)n_r * 32);
  A.n_l = nullptr; A.n_r = nullptr; A.n_l_host = n_l; A.n_r_host = n_r;
  A.cap = std::max(n_l, n_r);
  A.mbf = mbf; A.mb = mb;
  A.u_right = ar.alloc<float>(n_l);
  A.depth = ar.alloc<float>(n_l);
  A.sad = ar.alloc<int32_t>(n_l);
  A.n_matched = ar.alloc<int32_t>(1);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_stereo(A, 1, n_l, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  int32_t nm = 0;
  ORBM_CUDA(m, cudaMemcpyAsync(u_right, A.u_right, (size_t)n_l * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaMemcpyAsync(depth, A.depth, (size_t)n_l * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaMemcpyAsync(&nm, A.n_matched, 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  if (n_matched) *n_matched = nm;
  return ORBX_OK;
}

}  // extern "C"

namespace {
// the optional Tracking::SearchLocalPoints stage of the pipelined stereo call (host pointers)
struct TrackHost {
  const orbx_frustum* frustums;
  const orbx_local_map* maps;
  const int32_t* map_index;
  const uint8_t* occupied;
  const orbx_track_params* prm;
  int32_t *assign, *nmatches, *n_in_view;
};

constexpr int kRetryDirect = 0x7e7e0001;  // internal: a graph recording failed, run the call again on the direct path

int stereo_frames_body(orbm_matcher* m, orbx_extractor* left, orbx_extractor* right, int n_pairs,
                       const uint8_t* imgs_l, const uint8_t* imgs_r, int width, int height, int stride,
                       int64_t frame_stride, float mbf, float mb, orbx_kp* kps_l, uint8_t* desc_l,
                       int32_t* n_l, orbx_kp* kps_r, uint8_t* desc_r, int32_t* n_r, int cap, float* u_right,
                       float* depth, int32_t* n_matched, const TrackHost* trk, bool allow_graph) {
  if (!m || !left || !right || left == right) return mfail(m, ORBX_E_ARG, "bad argument");
  if (!imgs_l || !imgs_r || width <= 0 || height <= 0 || n_pairs <= 0) return mfail(m, ORBX_E_EMPTY, "empty image");
  if (stride < width || cap < 1 || !kps_l || !desc_l || !n_l || !kps_r || !desc_r || !n_r || !u_right || !depth ||
      !n_matched)
    return mfail(m, ORBX_E_ARG, "bad argument");
  if (left->device != m->device || right->device != m->device || left->max_batch != right->max_batch ||
      left->nlevels != right->nlevels)
    return mfail(m, ORBX_E_ARG, "extractors must share device, max_batch and level count with the matcher");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  int rc;
  for (orbx_extractor* ex : {left, right}) {
    if ((rc = api_ensure_plan(ex, width, height)) != 0 || (rc = api_ensure_out(ex, cap)) != 0)
      return mfail(m, rc, orbx_last_error(ex));
  }
  api_begin_call(left);
  api_begin_call(right);
  const int B = left->max_batch;
  for (int ln = 0; ln < kLanes; ln++) {
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 3 && e == cudaSuccess; k++) e = m->lane_buf[ln][k].reserve((size_t)B * cap * 4);
    if (e == cudaSuccess) e = m->lane_buf[ln][3].reserve((size_t)B * 4);
    if (e == cudaSuccess && m->lane_h_cap[ln] < B) {
      if (m->lane_h_nm[ln]) cudaFreeHost(m->lane_h_nm[ln]);
      m->lane_h_nm[ln] = nullptr;
      e = cudaHostAlloc(&m->lane_h_nm[ln], (size_t)B * 4, cudaHostAllocDefault);
      m->lane_h_cap[ln] = e == cudaSuccess ? B : 0;
    }
    if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
  }
  if (!m->up_stream) ORBM_CUDA(m, cudaStreamCreateWithFlags(&m->up_stream, cudaStreamNonBlocking));
  static const bool lane_uploads = getenv("ORBX_LANE_UPLOADS") != nullptr;  // A/B: uploads on the lanes' own streams
  const cudaStream_t up = m->up_stream;
  // ---- a call of ONE group that comes back unchanged is recorded as a CUDA graph on its second run and replayed from
  //      the third on (StereoGraphKey, orbm_handle.h): the single-pair call of an online front-end spends more host time
  //      issuing ~60 launches and ~20 copies than the GPU spends executing them. ORBX_GRAPH=0 keeps the direct path. ----
  enum { kDirect, kCapture, kReplay } mode = kDirect;
  StereoGraphSlot* slot = nullptr;
  static const bool graphs_on = [] {
    const char* e = getenv("ORBX_GRAPH");
    return !(e && e[0] == '0') && !getenv("ORBX_DEBUG_SKIP_KERNELS") && !getenv("ORBX_DEBUG_SKIP_H2D") &&
           !getenv("ORBX_TRACE") && !getenv("ORBX_SERIAL_EYES") && !getenv("ORBX_LANE_UPLOADS");
  }();
  if (allow_graph && graphs_on && !m->sgraph_off && n_pairs <= B && !left->profile && !right->profile &&
      (!trk || (trk->maps && trk->prm))) {
    StereoGraphKey key;
    memset(&key, 0, sizeof(key));
    const void* ptrs[] = {left, right, imgs_l, imgs_r, kps_l, desc_l, kps_r, desc_r, u_right, depth,
                          trk ? trk->frustums : nullptr, trk ? trk->maps->pos : nullptr, trk ? trk->maps->normal : nullptr,
                          trk ? trk->maps->min_dist : nullptr, trk ? trk->maps->max_dist : nullptr,
                          trk ? trk->maps->skip : nullptr, trk ? trk->maps->has_obs : nullptr,
                          trk ? trk->maps->desc : nullptr, trk ? trk->map_index : nullptr, trk ? trk->occupied : nullptr,
                          trk ? trk->assign : nullptr};
    static_assert(sizeof(ptrs) <= sizeof(key.ptr), "key");
    memcpy(key.ptr, ptrs, sizeof(ptrs));
    const long long ints[] = {n_pairs, width, height, stride, (long long)frame_stride, cap, trk ? 1 : 0,
                              trk ? trk->maps->m : 0, trk ? trk->maps->n_maps : 0, trk ? trk->prm->far_points : 0,
                              trk ? trk->prm->cand_per_frame : 0};
    static_assert(sizeof(ints) <= sizeof(key.iv), "key");
    memcpy(key.iv, ints, sizeof(ints));
    const float flts[] = {mbf, mb, trk ? trk->prm->viewing_cos_limit : 0.f, trk ? trk->prm->th : 0.f,
                          trk ? trk->prm->nnratio : 0.f, trk ? trk->prm->th_far : 0.f, trk ? trk->prm->min_x : 0.f,
                          trk ? trk->prm->min_y : 0.f, trk ? trk->prm->inv_w : 0.f, trk ? trk->prm->inv_h : 0.f};
    static_assert(sizeof(flts) <= sizeof(key.fv), "key");
    memcpy(key.fv, flts, sizeof(flts));
    key.gen = orbx::alloc_generation().load();
    StereoGraphSlot* victim = nullptr;  // a free slot, else the least recently used one
    for (StereoGraphSlot& sl : m->sgraph) {
      if (sl.used && memcmp(&sl.key, &key, sizeof(key)) == 0) slot = &sl;
      if (!victim || (victim->used && (!sl.used || sl.stamp < victim->stamp))) victim = &sl;
    }
    if (slot) {
      mode = slot->exec ? kReplay : kCapture;
    } else {  // first sight: run directly (every buffer gets its final size), remember the call
      if (victim->exec) cudaGraphExecDestroy(victim->exec);
      victim->exec = nullptr;
      victim->key = key;
      victim->used = true;
      slot = victim;
    }
    slot->stamp = ++m->sgraph_clock;
  }
  const cudaStream_t g_st = left->lane[0].stream, g_sr = right->lane[0].stream;
  // Ends an open recording on every way out of this function; the caller (stereo_frames_impl) sees sgraph_recording
  // still set, switches the handle to the direct path for good and runs the call again there.
  struct RecordingGuard {
    orbm_matcher* m;
    cudaStream_t s;
    ~RecordingGuard() {
      if (!m->sgraph_recording) return;
      cudaGraph_t g = nullptr;
      cudaStreamEndCapture(s, &g);
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
    }
  } recording_guard{m, g_st};
  if (mode == kCapture) {
    if (!m->sgraph_fork && cudaEventCreateWithFlags(&m->sgraph_fork, cudaEventDisableTiming) != cudaSuccess) mode = kDirect;
  }
  if (mode == kCapture) {
    // the left lane's stream is the origin; the right eye's stream and the upload stream fork from it here and join it
    // again through the events the direct path records anyway (lane.up, lane.done, track_map_ready)
    if (cudaStreamBeginCapture(g_st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      m->sgraph_off = true;
      mode = kDirect;
    } else {
      m->sgraph_recording = true;
      ORBM_CUDA(m, cudaEventRecord(m->sgraph_fork, g_st));
      ORBM_CUDA(m, cudaStreamWaitEvent(g_sr, m->sgraph_fork, 0));
      ORBM_CUDA(m, cudaStreamWaitEvent(up, m->sgraph_fork, 0));
    }
  }
  // ---- the tracking stage: local maps go to the device once per call, on the upload stream ----
  TrackArgs T0{};
  orbx_local_map dmap{};
  int32_t* d_map_index_all = nullptr;
  if (trk) {
    if (!trk->frustums || !trk->maps || !trk->prm || !trk->assign || !trk->nmatches || !trk->n_in_view)
      return mfail(m, ORBX_E_ARG, "bad argument");
    if ((rc = track_fill_params(m, left, trk->maps, trk->prm, left->lane[0].out_cap, &T0)) != 0) return rc;
    const orbx_local_map& hm = *trk->maps;
    const size_t tot = (size_t)hm.m * hm.n_maps;
    const size_t bytes[8] = {tot * 12, tot * 12, tot * 4, tot * 4, hm.skip ? tot : 0, tot, tot * 32,
                             trk->map_index ? (size_t)n_pairs * 4 : 0};
    const void* host[8] = {hm.pos, hm.normal, hm.min_dist, hm.max_dist, hm.skip, hm.has_obs, hm.desc, trk->map_index};
    if (!m->track_map_ready) ORBM_CUDA(m, cudaEventCreateWithFlags(&m->track_map_ready, cudaEventDisableTiming));
    for (int k = 0; k < 8; k++) {
      if (!bytes[k]) continue;
      cudaError_t e = m->track_map[k].reserve(bytes[k]);
      if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
      if (mode != kReplay)  // a replayed call has these copies in its graph
        ORBM_CUDA(m, cudaMemcpyAsync(m->track_map[k].p, host[k], bytes[k], cudaMemcpyHostToDevice, up));
    }
    if (mode != kReplay) ORBM_CUDA(m, cudaEventRecord(m->track_map_ready, up));
    dmap.m = hm.m;
    dmap.n_maps = hm.n_maps;
    dmap.pos = static_cast<const float*>(m->track_map[0].p);
    dmap.normal = static_cast<const float*>(m->track_map[1].p);
    dmap.min_dist = static_cast<const float*>(m->track_map[2].p);
    dmap.max_dist = static_cast<const float*>(m->track_map[3].p);
    dmap.skip = hm.skip ? static_cast<const uint8_t*>(m->track_map[4].p) : nullptr;
    dmap.has_obs = static_cast<const uint8_t*>(m->track_map[5].p);
    dmap.desc = static_cast<const uint8_t*>(m->track_map[6].p);
    d_map_index_all = trk->map_index ? static_cast<int32_t*>(m->track_map[7].p) : nullptr;
    // without an explicit array pair p uses map p % n_maps: the kernels compute it from TrackArgs::map_f0
    track_set_map(&dmap, &T0);
    for (int ln = 0; ln < kLanes; ln++) {
      if (m->lane_h_track_cap[ln] < B) {
        if (m->lane_h_track[ln]) cudaFreeHost(m->lane_h_track[ln]);
        m->lane_h_track[ln] = nullptr;
        cudaError_t e = cudaHostAlloc(&m->lane_h_track[ln], (size_t)3 * B * 4, cudaHostAllocDefault);
        m->lane_h_track_cap[ln] = e == cudaSuccess ? B : 0;
        if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
      }
    }
  }
  int first_err = ORBX_OK;
  int pending_f0[kLanes], pending_nb[kLanes];
  for (int i = 0; i < kLanes; i++) pending_nb[i] = 0;
  auto retire = [&](int ln, bool synced = false) -> int {
    if (pending_nb[ln] == 0) return ORBX_OK;
    if (!synced) ORBM_CUDA(m, cudaEventSynchronize(left->lane[ln].done));
    for (int f = 0; f < pending_nb[ln]; f++) {
      const int g = pending_f0[ln] + f;
      n_l[g] = left->lane[ln].h_small[f];
      n_r[g] = right->lane[ln].h_small[f];
      n_matched[g] = m->lane_h_nm[ln][f];
      bool bad = left->lane[ln].h_small[2 * B + f] != 0 || right->lane[ln].h_small[2 * B + f] != 0 ||
                 n_l[g] > cap || n_r[g] > cap;
      if (trk) {
        trk->nmatches[g] = m->lane_h_track[ln][f];
        trk->n_in_view[g] = m->lane_h_track[ln][B + f];
        bad = bad || m->lane_h_track[ln][2 * B + f] != 0;
      }
      if (bad && first_err == ORBX_OK) first_err = ORBX_E_CAPACITY;
    }
    pending_nb[ln] = 0;
    return ORBX_OK;
  };
  // Group sizes: full groups of B pairs, except that a long call ramps up (B/4, B/2) and down (B/2, B/4): nothing can
  // run before the first group's images have crossed PCIe and nothing overlaps the last group's kernels, so those two
  // are kept short (fill 1.7 -> 0.4 ms at B = 128 and 752x480).
  std::vector<int> sizes;
  if (n_pairs >= 3 * B && B >= 8) {
    const int mid = n_pairs - (B / 4 + B / 2) * 2;
    sizes.push_back(B / 4);
    sizes.push_back(B / 2);
    for (int k = 0; k < mid / B; k++) sizes.push_back(B);
    if (mid % B) sizes.push_back(mid % B);
    sizes.push_back(B / 2);
    sizes.push_back(B / 4);
  } else {
    for (int f0 = 0; f0 < n_pairs; f0 += B) sizes.push_back(std::min(B, n_pairs - f0));
  }
  // ORBX_TRACE=1: host time spent waiting for a lane vs issuing work, per call (stderr)
  static const bool trace = getenv("ORBX_TRACE") != nullptr;
  static const bool timeline = trace && atoi(getenv("ORBX_TRACE")) >= 2;  // per-group GPU timeline (CUDA events)
  std::vector<cudaEvent_t> tev;  // 9 per group, see the print below
  auto tl = [&](int which, cudaStream_t s) {
    if (!timeline) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    tev.push_back(e);
    (void)which;
    cudaEventRecord(e, s);
  };
  auto tl_slot = [&]() -> cudaEvent_t {
    if (!timeline) return nullptr;
    cudaEvent_t e;
    cudaEventCreate(&e);
    tev.push_back(e);
    return e;
  };
  auto finish_recorded = [&]() -> int {  // launch the recorded call, wait for it, hand out the pinned result words
    ORBM_CUDA(m, cudaGraphLaunch(slot->exec, g_st));
    ORBM_CUDA(m, cudaStreamSynchronize(g_st));
    m->sgraph_launches++;
    for (orbx_extractor* ex : {left, right}) {  // what run_pipeline leaves behind on the host side
      OrbxLane& L = ex->lane[0];
      L.last_fs = ex == left ? slot->fs_l : slot->fs_r;
      L.last_frames = n_pairs;
      L.last_f0 = 0;
      L.last_call = ex->call_id;
      ex->last_lane = 0;
    }
    pending_f0[0] = 0;
    pending_nb[0] = n_pairs;
    int r = retire(0, true);
    if (r) return r;
    if (first_err) return mfail(m, first_err, "output capacity (or the candidate list) too small for at least one frame");
    return ORBX_OK;
  };
  if (mode == kReplay) return finish_recorded();
  double t_wait = 0, t_issue = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  auto issue_group = [&](size_t gi, int f0, int group) -> int {
    const int nb = sizes[gi];
    static const int n_lanes = getenv("ORBX_LANES") ? std::max(1, std::min((int)kLanes, atoi(getenv("ORBX_LANES")))) : (int)kLanes;
    const int ln = group % n_lanes;
    const double tw0 = trace ? now() : 0;
    if ((rc = retire(ln)) != 0) return rc;
    const double tw1 = trace ? now() : 0;
    t_wait += tw1 - tw0;
    struct IssueTimer {
      double& acc; double t0; bool on; decltype(now)& clk;
      ~IssueTimer() { if (on) acc += clk() - t0; }
    } issue_timer{t_issue, tw1, trace, now};
    // The two eyes run on their own streams, like the two std::threads of the reference's stereo constructor
    // (src/Frame.cc:200-203): the latency-bound stages of one eye (quadtree, small pyramid levels) overlap the
    // ALU-bound stages of the other, and with kLanes groups in flight the copy engines stay busy too.
    static const bool serial_eyes = getenv("ORBX_SERIAL_EYES") != nullptr;  // A/B experiment: both eyes on one stream
    cudaStream_t st = left->lane[ln].stream, sr = serial_eyes ? st : right->lane[ln].stream;
    // The tracking stage's small per-group inputs (poses, occupancy) cross PCIe FIRST: enqueued behind the stereo
    // kernels they would sit in the copy engine's queue behind the next lanes' image uploads (measured: the call took
    // 20.2 ms against 13.0 ms without the image uploads and 12.2 ms for the uploads alone)
    orbx_frustum* d_fr = nullptr;
    int32_t* d_assign = nullptr;
    int32_t* d_words = nullptr;
    uint8_t* d_occ = nullptr;
    if (trk) {
      const int dcap0 = left->lane[ln].out_cap;
      DevBuf* tb = m->track_buf[ln];
      cudaError_t e = tb[9].reserve((size_t)B * sizeof(orbx_frustum));
      if (e == cudaSuccess) e = tb[10].reserve((size_t)B * dcap0 * 4 + (size_t)B * 3 * 4);
      if (e == cudaSuccess && trk->occupied) e = tb[11].reserve((size_t)B * dcap0);
      if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
      d_fr = static_cast<orbx_frustum*>(tb[9].p);
      d_assign = static_cast<int32_t*>(tb[10].p);
      d_words = d_assign + (size_t)B * dcap0;  // nmatches[B] | n_in_view[B] | status[B]
      const cudaStream_t cs = lane_uploads ? st : up;
      ORBM_CUDA(m, cudaMemcpyAsync(d_fr, trk->frustums + f0, (size_t)nb * sizeof(orbx_frustum), cudaMemcpyHostToDevice, cs));
      if (trk->occupied) {
        d_occ = static_cast<uint8_t*>(tb[11].p);
        if (cap == dcap0) {
          ORBM_CUDA(m, cudaMemcpyAsync(d_occ, trk->occupied + (size_t)f0 * cap, (size_t)nb * cap, cudaMemcpyHostToDevice, cs));
        } else {
          ORBM_CUDA(m, cudaMemsetAsync(d_occ, 0, (size_t)nb * dcap0, cs));
          ORBM_CUDA(m, cudaMemcpy2DAsync(d_occ, dcap0, trk->occupied + (size_t)f0 * cap, cap, std::min(cap, dcap0), nb,
                                         cudaMemcpyHostToDevice, cs));
        }
      }
      if (!lane_uploads) {
        if (!m->up_small[ln]) ORBM_CUDA(m, cudaEventCreateWithFlags(&m->up_small[ln], cudaEventDisableTiming));
        ORBM_CUDA(m, cudaEventRecord(m->up_small[ln], up));
        ORBM_CUDA(m, cudaStreamWaitEvent(st, m->up_small[ln], 0));
      }
    }
    tl(0, st);
    orbx::g_trace_after_h2d = tl_slot();
    if ((rc = api_upload_and_run(left, ln, imgs_l + (int64_t)f0 * frame_stride, nb, width, height, stride,
                                 frame_stride, 0, 0, st, f0, lane_uploads ? nullptr : up)) != 0)
      return mfail(m, rc, orbx_last_error(left));
    tl(2, st);
    tl(5, sr);
    orbx::g_trace_after_h2d = tl_slot();
    if ((rc = api_upload_and_run(right, ln, imgs_r + (int64_t)f0 * frame_stride, nb, width, height, stride,
                                 frame_stride, 0, 0, sr, f0, lane_uploads ? nullptr : up)) != 0)
      return mfail(m, rc, orbx_last_error(right));
    orbx::g_trace_after_h2d = nullptr;
    tl(7, sr);
    // the right eye's descriptors can go home while the stereo matcher runs
    if ((rc = api_download(right, ln, nb, kps_r + (int64_t)f0 * cap, desc_r + (int64_t)f0 * cap * 32, cap, sr)) != 0)
      return mfail(m, rc, orbx_last_error(right));
    tl(8, sr);
    ORBM_CUDA(m, cudaEventRecord(right->lane[ln].done, sr));
    ORBM_CUDA(m, cudaStreamWaitEvent(st, right->lane[ln].done, 0));
    // ... and ComputeStereoMatches (:223) on the outputs still resident in the lanes
    StereoArgs A;
    if ((rc = stereo_args_common(m, left, right, &A, ln, ln)) != 0) return rc;
    const OrbxLane &LL = left->lane[ln], &RL = right->lane[ln];
    const int dcap = LL.out_cap;
    if (RL.out_cap != dcap) return mfail(m, ORBX_E_ARG, "extractor output capacities differ");
    A.frame0 = 0;
    A.kps_l = LL.d_kps; A.desc_l = LL.d_desc; A.n_l = LL.d_n;
    A.kps_r = RL.d_kps; A.desc_r = RL.d_desc; A.n_r = RL.d_n;
    A.cap = dcap; A.mbf = mbf; A.mb = mb;
    // the lane buffers were sized for `cap` rows per pair; rows are addressed with the device capacity
    for (int k = 0; k < 3; k++) {
      cudaError_t e = m->lane_buf[ln][k].reserve((size_t)B * dcap * 4);
      if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
    }
    A.u_right = reinterpret_cast<float*>(m->lane_buf[ln][0].p);
    A.depth = reinterpret_cast<float*>(m->lane_buf[ln][1].p);
    A.sad = reinterpret_cast<int32_t*>(m->lane_buf[ln][2].p);
    A.n_matched = reinterpret_cast<int32_t*>(m->lane_buf[ln][3].p);
    static const bool skip_kernels = getenv("ORBX_DEBUG_SKIP_KERNELS") != nullptr;  // timing experiment only
    if (!skip_kernels) launch_stereo(A, nb, dcap, st);
    ORBM_CUDA(m, cudaGetLastError());
    tl(3, st);
    if ((rc = api_download(left, ln, nb, kps_l + (int64_t)f0 * cap, desc_l + (int64_t)f0 * cap * 32, cap, st)) != 0)
      return mfail(m, rc, orbx_last_error(left));
    if (cap == dcap) {
      ORBM_CUDA(m, cudaMemcpyAsync(u_right + (int64_t)f0 * cap, A.u_right, (size_t)nb * cap * 4,
                                   cudaMemcpyDeviceToHost, st));
      ORBM_CUDA(m, cudaMemcpyAsync(depth + (int64_t)f0 * cap, A.depth, (size_t)nb * cap * 4, cudaMemcpyDeviceToHost,
                                   st));
    } else {
      const int rows = std::min(cap, dcap);
      ORBM_CUDA(m, cudaMemcpy2DAsync(u_right + (int64_t)f0 * cap, (size_t)cap * 4, A.u_right, (size_t)dcap * 4,
                                     (size_t)rows * 4, nb, cudaMemcpyDeviceToHost, st));
      ORBM_CUDA(m, cudaMemcpy2DAsync(depth + (int64_t)f0 * cap, (size_t)cap * 4, A.depth, (size_t)dcap * 4,
                                     (size_t)rows * 4, nb, cudaMemcpyDeviceToHost, st));
    }
    ORBM_CUDA(m, cudaMemcpyAsync(m->lane_h_nm[ln], A.n_matched, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
    if (trk) {
      // Tracking::SearchLocalPoints on the left frames of the group, which are still resident in the lane
      TrackArgs T = T0;
      T.n_frames = nb;
      T.kps = LL.d_kps;
      T.desc = LL.d_desc;
      T.n = LL.d_n;
      T.u_right = A.u_right;
      if ((rc = track_prepare(m, ln, &T)) != 0) return rc;
      T.frustums = d_fr;
      T.map_index = nullptr;
      if (d_map_index_all) T.map_index = d_map_index_all + f0;
      T.map_f0 = f0;  // no explicit array: pair f0 + f uses map (f0 + f) % n_maps
      T.occupied = trk->occupied ? d_occ : nullptr;
      T.assign = d_assign;
      T.nmatches = d_words;
      T.n_in_view = d_words + B;
      T.status = d_words + 2 * B;
      ORBM_CUDA(m, cudaStreamWaitEvent(st, m->track_map_ready, 0));
      if (!skip_kernels) {
        launch_frustum_batch(T, st);
        launch_track_search(T, st);
      }
      ORBM_CUDA(m, cudaGetLastError());
      if (cap == dcap) {
        ORBM_CUDA(m, cudaMemcpyAsync(trk->assign + (int64_t)f0 * cap, d_assign, (size_t)nb * cap * 4,
                                     cudaMemcpyDeviceToHost, st));
      } else {
        ORBM_CUDA(m, cudaMemcpy2DAsync(trk->assign + (int64_t)f0 * cap, (size_t)cap * 4, d_assign, (size_t)dcap * 4,
                                       (size_t)std::min(cap, dcap) * 4, nb, cudaMemcpyDeviceToHost, st));
      }
      for (int k = 0; k < 3; k++)
        ORBM_CUDA(m, cudaMemcpyAsync(m->lane_h_track[ln] + (size_t)k * B, d_words + (size_t)k * B, (size_t)nb * 4,
                                     cudaMemcpyDeviceToHost, st));
    }
    tl(4, st);
    ORBM_CUDA(m, cudaEventRecord(left->lane[ln].done, st));
    pending_f0[ln] = f0;
    pending_nb[ln] = nb;
    return ORBX_OK;
  };
  {
    int group = 0, f0 = 0;
    for (size_t gi = 0; gi < sizes.size(); f0 += sizes[gi], gi++, group++)
      if ((rc = issue_group(gi, f0, group)) != 0) return rc;  // (an open recording is ended by recording_guard)
  }
  if (mode == kCapture) {
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(g_st, &g);
    cudaError_t ge = ce;
    if (ce == cudaSuccess && g) ge = cudaGraphInstantiate(&slot->exec, g, 0);
    if (g) cudaGraphDestroy(g);
    if (ge != cudaSuccess || !slot->exec) {
      slot->exec = nullptr;
      cudaGetLastError();
      m->sgraph_recording = false;  // the capture is closed; tell the caller to go direct
      m->sgraph_off = true;
      return kRetryDirect;
    }
    m->sgraph_recording = false;
    slot->fs_l = left->lane[0].last_fs;
    slot->fs_r = right->lane[0].last_fs;
    pending_nb[0] = 0;  // nothing ran yet: the recorded work is launched now
    return finish_recorded();
  }
  const double td0 = trace ? now() : 0;
  for (int ln = 0; ln < kLanes; ln++)
    if ((rc = retire(ln)) != 0) return rc;
  if (trace)
    fprintf(stderr, "orbm_stereo_frames_batch: %zu groups, host issue %.2f ms, wait for lanes %.2f ms, drain %.2f ms\n",
            sizes.size(), 1e3 * t_issue, 1e3 * t_wait, 1e3 * (now() - td0));
  if (timeline && tev.size() == 9 * sizes.size() && sizes.size() > 2) {
    // per group, ms since the first event. Recording order: L.start, L.h2d (images up), L.extr (extractor done), R.start,
    // R.h2d, R.extr, R.d2h (right results home), L.stereo (stereo matcher done), L.end (tracking + all results home)
    fprintf(stderr, "group pairs |  L.start   L.h2d  L.extr L.stereo   L.end |  R.start   R.h2d  R.extr   R.d2h\n");
    for (size_t g = 0; g < sizes.size(); g++) {
      float t[9];
      for (int k = 0; k < 9; k++) cudaEventElapsedTime(&t[k], tev[0], tev[9 * g + k]);
      fprintf(stderr, "%5zu %5d | %8.2f %7.2f %7.2f %8.2f %7.2f | %8.2f %7.2f %7.2f %7.2f\n", g, sizes[g], t[0], t[1], t[2], t[7],
              t[8], t[3], t[4], t[5], t[6]);
    }
  }
  for (cudaEvent_t e : tev) cudaEventDestroy(e);
  if (first_err) return mfail(m, first_err, "output capacity (or the candidate list) too small for at least one frame");
  return ORBX_OK;
}

int stereo_frames_impl(orbm_matcher* m, orbx_extractor* left, orbx_extractor* right, int n_pairs,
                       const uint8_t* imgs_l, const uint8_t* imgs_r, int width, int height, int stride,
                       int64_t frame_stride, float mbf, float mb, orbx_kp* kps_l, uint8_t* desc_l,
                       int32_t* n_l, orbx_kp* kps_r, uint8_t* desc_r, int32_t* n_r, int cap, float* u_right,
                       float* depth, int32_t* n_matched, const TrackHost* trk) {
  int rc = stereo_frames_body(m, left, right, n_pairs, imgs_l, imgs_r, width, height, stride, frame_stride, mbf, mb, kps_l,
                              desc_l, n_l, kps_r, desc_r, n_r, cap, u_right, depth, n_matched, trk, true);
  if (m && m->sgraph_recording) {  // left early with a recording open (closed by its guard)
    m->sgraph_recording = false;
    m->sgraph_off = true;
    rc = kRetryDirect;
  }
  if (rc == kRetryDirect)  // the recording met something a stream capture cannot hold: the handle stays direct from now on
    rc = stereo_frames_body(m, left, right, n_pairs, imgs_l, imgs_r, width, height, stride, frame_stride, mbf, mb, kps_l,
                            desc_l, n_l, kps_r, desc_r, n_r, cap, u_right, depth, n_matched, trk, false);
  return rc;
}
}  // namespace

extern "C" {

/* Development aid (not in include/orbm.h): how many small stereo calls ran as a recorded CUDA graph on this handle. */
int orbm_debug_graph_launches(const orbm_matcher* m) { return m ? (int)m->sgraph_launches : -1; }

int orbm_stereo_frames_batch(orbm_matcher* m, orbx_extractor* left, orbx_extractor* right, int n_pairs,
                             const uint8_t* imgs_l, const uint8_t* imgs_r, int width, int height, int stride,
                             int64_t frame_stride, float mbf, float mb, orbx_kp* kps_l, uint8_t* desc_l,
                             int32_t* n_l, orbx_kp* kps_r, uint8_t* desc_r, int32_t* n_r, int cap, float* u_right,
                             float* depth, int32_t* n_matched) {
  return stereo_frames_impl(m, left, right, n_pairs, imgs_l, imgs_r, width, height, stride, frame_stride, mbf, mb, kps_l,
                            desc_l, n_l, kps_r, desc_r, n_r, cap, u_right, depth, n_matched, nullptr);
}

int orbm_stereo_track_frames_batch(orbm_matcher* m, orbx_extractor* left, orbx_extractor* right, int n_pairs,
                                   const uint8_t* imgs_l, const uint8_t* imgs_r, int width, int height, int stride,
                                   int64_t frame_stride, float mbf, float mb, const orbx_frustum* frustums,
                                   const orbx_local_map* maps, const int32_t* map_index, const uint8_t* occupied,
                                   const orbx_track_params* prm, orbx_kp* kps_l, uint8_t* desc_l, int32_t* n_l,
                                   orbx_kp* kps_r, uint8_t* desc_r, int32_t* n_r, int cap, float* u_right, float* depth,
                                   int32_t* n_matched, int32_t* assign, int32_t* nmatches, int32_t* n_in_view) {
  const TrackHost trk{frustums, maps, map_index, occupied, prm, assign, nmatches, n_in_view};
  return stereo_frames_impl(m, left, right, n_pairs, imgs_l, imgs_r, width, height, stride, frame_stride, mbf, mb, kps_l,
                            desc_l, n_l, kps_r, desc_r, n_r, cap, u_right, depth, n_matched, &trk);
}

int orbm_stereo_track_frames_batch_multi(int n_devices, orbm_matcher* const* m, orbx_extractor* const* left,
                                         orbx_extractor* const* right, int n_pairs, const uint8_t* imgs_l,
                                         const uint8_t* imgs_r, int width, int height, int stride, int64_t frame_stride,
                                         float mbf, float mb, const orbx_frustum* frustums, const orbx_local_map* maps,
                                         const int32_t* map_index, const uint8_t* occupied, const orbx_track_params* prm,
                                         orbx_kp* kps_l, uint8_t* desc_l, int32_t* n_l, orbx_kp* kps_r, uint8_t* desc_r,
                                         int32_t* n_r, int cap, float* u_right, float* depth, int32_t* n_matched,
                                         int32_t* assign, int32_t* nmatches, int32_t* n_in_view) {
  if (n_devices < 1 || !m || !left || !right) return ORBX_E_ARG;
  for (int d = 0; d < n_devices; d++)
    if (!m[d] || !left[d] || !right[d]) return ORBX_E_ARG;
  if (n_pairs <= 0) return mfail(m[0], ORBX_E_EMPTY, "empty image");
  const bool track = frustums != nullptr;
  // the default map rule "pair p uses map p % n_maps" is global: a block that does not start at 0 needs it spelled out
  std::vector<int32_t> idx;
  if (track && !map_index && maps && maps->n_maps > 1) {
    idx.resize(n_pairs);
    for (int p = 0; p < n_pairs; p++) idx[p] = p % maps->n_maps;
    map_index = idx.data();
  }
  std::vector<int> rc(n_devices, ORBX_OK);
  std::vector<std::thread> th;
  const int base = n_pairs / n_devices, extra = n_pairs % n_devices;
  int f0 = 0;
  for (int d = 0; d < n_devices; d++) {
    const int nb = base + (d < extra ? 1 : 0);
    if (nb > 0)
      th.emplace_back([=, &rc] {
        const int64_t o = f0, oc = (int64_t)f0 * cap;
        TrackHost trk{track ? frustums + o : nullptr, maps, map_index ? map_index + o : nullptr,
                      occupied ? occupied + oc : nullptr, prm, assign ? assign + oc : nullptr,
                      nmatches ? nmatches + o : nullptr, n_in_view ? n_in_view + o : nullptr};
        rc[d] = stereo_frames_impl(m[d], left[d], right[d], nb, imgs_l + o * frame_stride, imgs_r + o * frame_stride, width,
                                   height, stride, frame_stride, mbf, mb, kps_l + oc, desc_l + oc * 32, n_l + o, kps_r + oc,
                                   desc_r + oc * 32, n_r + o, cap, u_right + oc, depth + oc, n_matched + o,
                                   track ? &trk : nullptr);
      });
    f0 += nb;
  }
  for (auto& t : th) t.join();
  for (int d = 0; d < n_devices; d++)
    if (rc[d] != ORBX_OK) return rc[d];
  return ORBX_OK;
}

namespace {
// The part of SearchByProjection(Frame&, vector<MapPoint*>) that does not depend on where the frame lives: the scalar
// prologue of the reference loop (:55-74) — which points take part, the search radius
// r = RadiusByViewingCos(viewCos) * th * scale[level] (:65-67, :223-228), the level window [L-1, L] — then the search.
int search_map_common(orbm_matcher* m, Arena& ar, const DevFrame& F, const float* scale_factors, int n_levels,
                      bool has_u_right, const orbx_mappoints* mps, float th, float nnratio, int far_points,
                      float th_far, int n, int32_t* assign, int32_t* nmatches) {
  const int M = mps->m;
  std::vector<uint8_t> active(std::max(M, 1));
  std::vector<float> radius(std::max(M, 1));
  std::vector<int32_t> minl(std::max(M, 1)), maxl(std::max(M, 1));
  const bool bFactor = th != 1.0;
  for (int i = 0; i < M; i++) {
    bool on = mps->track_in_view[i] != 0;
    if (on && far_points && mps->depth[i] > th_far) on = false;
    const int level = mps->level[i];
    if (on && (level < 0 || level >= n_levels)) on = false;
    active[i] = on;
    float r = ((double)mps->view_cos[i] > 0.998) ? 2.5f : 4.0f;
    if (bFactor) r *= th;
    radius[i] = on ? r * scale_factors[level] : 0.f;
    minl[i] = level - 1;
    maxl[i] = level;
  }
  DevQueries Q{};
  Q.m = M;
  Q.active = ar.upload(active.data(), M);
  Q.u = ar.upload(mps->proj_x, M);
  Q.v = ar.upload(mps->proj_y, M);
  Q.radius = ar.upload(radius.data(), M);
  Q.min_level = ar.upload(minl.data(), M);
  Q.max_level = ar.upload(maxl.data(), M);
  Q.u_right = has_u_right ? ar.upload(mps->proj_xr, M) : (ar.alloc<float>(1), nullptr);
  Q.desc = ar.upload(mps->desc, (size_t)M * 32);
  ResolveArgs R{};
  R.mode = 0;
  R.nnratio = nnratio;
  R.max_dist = ORBM_TH_HIGH_I;
  R.check_orientation = 0;
  R.has_obs = ar.upload(mps->has_obs, M);
  R.angle = nullptr;
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  return run_search(m, ar, F, Q, R, n, assign, nmatches);
}
}  // namespace

int orbm_search_by_projection_map(orbm_matcher* m, const orbx_frame_view* f, const orbx_mappoints* mps, float th,
                                  float nnratio, int far_points, float th_far, int32_t* assign, int32_t* nmatches) {
  if (!m || !f || !mps || f->n < 0 || mps->m < 0 || (f->n > 0 && !assign)) return mfail(m, ORBX_E_ARG, "bad argument");
  if (nmatches) *nmatches = 0;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  const DevFrame F = upload_frame(ar, f);
  return search_map_common(m, ar, F, f->scale_factors, f->n_levels, f->u_right != nullptr, mps, th, nnratio, far_points,
                           th_far, f->n, assign, nmatches);
}

int orbm_search_by_projection_map_fisheye(orbm_matcher* m, const orbx_fisheye_view* f, const orbx_mappoints* mps,
                                          const orbx_mappoints_right* mr, float th, float nnratio, int far_points,
                                          float th_far, int32_t* assign, int32_t* nmatches) {
  if (!m || !f || !mps || !mr || f->n_left < 0 || f->n_right < 0 || mps->m < 0 || (f->n_left + f->n_right > 0 && !assign) ||
      !f->left_to_right || !f->right_to_left)
    return mfail(m, ORBX_E_ARG, "bad argument");
  if (nmatches) *nmatches = 0;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  const int NL = f->n_left, NR = f->n_right, N = NL + NR, M = mps->m, cells = ORBX_GRID_COLS * ORBX_GRID_ROWS;
  if (N == 0) return ORBX_OK;
  Arena ar(m);
  cudaStream_t st = m->stream;
  // the enumeration must not drop occupied keypoints: a partner write can re-open a slot (see k_search_resolve_fisheye)
  std::vector<uint8_t> zeros(std::max(N, 1), 0);
  const uint8_t* d_zero = ar.upload(zeros.data(), N);
  const uint8_t* d_desc = ar.upload(f->desc, (size_t)N * 32);
  const float* d_sf = ar.upload(f->scale_factors, f->n_levels);
  DevFrame F[2];
  for (int side = 0; side < 2; side++) {
    const orbx_grid& g = side ? f->grid_right : f->grid_left;
    DevFrame& D = F[side];
    D = DevFrame{};
    D.n = side ? NR : NL;
    D.n_levels = f->n_levels;
    D.kps = ar.upload(side ? f->kps_right : f->kps_left, D.n);
    D.desc = d_desc + (side ? (size_t)NL * 32 : 0);
    D.u_right = nullptr;
    D.occupied = d_zero;
    D.cell_offsets = ar.upload(g.cell_offsets, cells + 1);
    D.cell_items = ar.upload(g.cell_items, (size_t)g.cell_offsets[cells]);
    D.min_x = f->grid_left.min_x;
    D.min_y = f->grid_left.min_y;
    D.inv_w = f->grid_left.inv_w;
    D.inv_h = f->grid_left.inv_h;
    D.scale_factors = d_sf;
  }
  // the scalar prologue of both loop bodies (:55-74, :148-160): who searches where, with which radius
  std::vector<uint8_t> act[2] = {std::vector<uint8_t>(std::max(M, 1)), std::vector<uint8_t>(std::max(M, 1))};
  std::vector<float> radius[2] = {std::vector<float>(std::max(M, 1)), std::vector<float>(std::max(M, 1))};
  std::vector<int32_t> minl[2] = {std::vector<int32_t>(std::max(M, 1)), std::vector<int32_t>(std::max(M, 1))};
  std::vector<int32_t> maxl[2] = {std::vector<int32_t>(std::max(M, 1)), std::vector<int32_t>(std::max(M, 1))};
  const bool bFactor = th != 1.0;
  for (int i = 0; i < M; i++) {
    const bool inL = mps->track_in_view[i] != 0, inR = mr->track_in_view_r[i] != 0;
    const bool gone = (!inL && !inR) || (far_points && mps->depth[i] > th_far);
    for (int side = 0; side < 2; side++) {
      const int level = side ? mr->level_r[i] : mps->level[i];
      bool on = !gone && (side ? inR : inL);
      if (on && side == 1 && level == -1) on = false;  // :150
      if (on && (level < 0 || level >= f->n_levels)) on = false;
      float r = ((double)(side ? mr->view_cos_r[i] : mps->view_cos[i]) > 0.998) ? 2.5f : 4.0f;
      if (side == 0 && bFactor) r *= th;  // the right-camera twin has no th factor (:152)
      act[side][i] = on;
      radius[side][i] = on ? r * f->scale_factors[level] : 0.f;
      minl[side][i] = level - 1;
      maxl[side][i] = level;
    }
  }
  const uint8_t* d_q_desc = ar.upload(mps->desc, (size_t)M * 32);
  FisheyeResolveArgs R{};
  R.m = M;
  R.n_left = NL;
  R.n_right = NR;
  R.nnratio = nnratio;
  R.has_obs = ar.upload(mps->has_obs, M);
  R.occupied = f->occupied ? ar.upload(f->occupied, N) : d_zero;
  R.left_to_right = ar.upload(f->left_to_right, NL);
  R.right_to_left = ar.upload(f->right_to_left, NR);
  R.assign = ar.alloc<int32_t>(N);
  R.nmatches = ar.alloc<int32_t>(1);
  DevQueries Q[2];
  SearchScratch S[2];
  int32_t* d_total[2];
  for (int side = 0; side < 2; side++) {
    Q[side] = DevQueries{};
    Q[side].m = M;
    Q[side].active = ar.upload(act[side].data(), M);
    Q[side].u = ar.upload(side ? mr->proj_x_r : mps->proj_x, M);
    Q[side].v = ar.upload(side ? mr->proj_y_r : mps->proj_y, M);
    Q[side].radius = ar.upload(radius[side].data(), M);
    Q[side].min_level = ar.upload(minl[side].data(), M);
    Q[side].max_level = ar.upload(maxl[side].data(), M);
    Q[side].u_right = nullptr;
    Q[side].desc = d_q_desc;
    S[side] = SearchScratch{};
    S[side].counts = ar.alloc<int32_t>((size_t)M + 1);
    S[side].pre = ar.alloc<int4>(M);
    d_total[side] = ar.alloc<int32_t>(1);
  }
  R.in_left = Q[0].active;
  R.in_right = Q[1].active;
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  int32_t total[2] = {0, 0};
  for (int side = 0; side < 2; side++) {
    if (M > 0) {
      launch_search_count(F[side], Q[side], S[side].counts, st);
      launch_scan(S[side].counts, M, d_total[side], st);
      ORBM_CUDA(m, cudaMemcpyAsync(&total[side], d_total[side], 4, cudaMemcpyDeviceToHost, st));
    } else {
      ORBM_CUDA(m, cudaMemsetAsync(S[side].counts, 0, 4, st));
    }
  }
  ORBM_CUDA(m, cudaStreamSynchronize(st));
  for (int side = 0; side < 2; side++) {
    S[side].cand_idx = ar.alloc<int32_t>(total[side]);
    S[side].cand_dist = ar.alloc<int32_t>(total[side]);
    S[side].cap_total = total[side];
  }
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  for (int side = 0; side < 2; side++) launch_search_fill(F[side], Q[side], S[side], st);
  R.left = S[0];
  R.right = S[1];
  launch_search_resolve_fisheye(R, st);
  ORBM_CUDA(m, cudaGetLastError());
  int32_t nm = 0;
  ORBM_CUDA(m, cudaMemcpyAsync(assign, R.assign, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(&nm, R.nmatches, 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaStreamSynchronize(st));
  if (nmatches) *nmatches = nm;
  return ORBX_OK;
}

int orbm_assign_features_to_grid(orbm_matcher* m, const orbx_kp* kps, int n, float min_x, float min_y, float inv_w,
                                 float inv_h, int32_t* cell_offsets, int32_t* cell_items) {
  if (!m || n < 0 || (n > 0 && (!kps || !cell_items)) || !cell_offsets) return mfail(m, ORBX_E_ARG, "bad argument");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  const int cells = ORBX_GRID_COLS * ORBX_GRID_ROWS;
  Arena ar(m);
  const orbx_kp* d_kps = ar.upload(kps, n);
  int32_t* d_off = ar.alloc<int32_t>(cells + 1);
  int32_t* d_items = ar.alloc<int32_t>(n);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_build_grid(d_kps, nullptr, n, 0, 1, min_x, min_y, inv_w, inv_h, d_off, d_items, 0, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  ORBM_CUDA(m, cudaMemcpyAsync(cell_offsets, d_off, (size_t)(cells + 1) * 4, cudaMemcpyDeviceToHost, m->stream));
  if (n > 0) ORBM_CUDA(m, cudaMemcpyAsync(cell_items, d_items, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  return ORBX_OK;
}

namespace {
// DevFrame of frame `frame` of the extractor's most recent host-facing call: keypoints / descriptors where the
// extractor left them, mvScaleFactors from its plan, the 64x48 grid built on the device (Frame::AssignFeaturesToGrid).
int resident_frame(orbm_matcher* m, Arena& ar, const orbx_extractor* ex, int frame, int n, const float* u_right,
                   const uint8_t* occupied, float min_x, float min_y, float inv_w, float inv_h, DevFrame* out,
                   float* sf) {
  if (!ex->planned || ex->device != m->device) return mfail(m, ORBX_E_ARG, "extractor has not run on this device");
  int fl = 0, lf = 0;
  if (api_find_frame(ex, frame, &fl, &lf) != ORBX_OK)
    return mfail(m, ORBX_E_ARG, "frame is not resident in the extractor's last call (only its last kLanes groups are)");
  const OrbxLane& L = ex->lane[fl];
  frame = lf;
  if (!L.d_kps || n > L.out_cap) return mfail(m, ORBX_E_ARG, "frame has no device outputs (host-facing calls only)");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  const int cells = ORBX_GRID_COLS * ORBX_GRID_ROWS;
  const Plan& P = ex->plan;
  for (int l = 0; l < P.nlevels; l++) sf[l] = P.lv[l].scale;  // mvScaleFactors (src/ORBextractor.cc:418-425)
  DevFrame F{};
  F.n = n;
  F.n_levels = P.nlevels;
  F.kps = L.d_kps + (int64_t)frame * L.out_cap;  // mvKeysUn == mvKeys for an undistorted camera (src/Frame.cc:562-571)
  F.desc = L.d_desc + (int64_t)frame * L.out_cap * 32;
  F.u_right = u_right ? ar.upload(u_right, n) : (ar.alloc<float>(1), nullptr);
  uint8_t* d_occ = ar.alloc<uint8_t>(n);
  int32_t* d_off = ar.alloc<int32_t>(cells + 1);
  int32_t* d_items = ar.alloc<int32_t>(n);
  F.scale_factors = ar.upload(sf, P.nlevels);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  if (n > 0) {
    if (occupied) ORBM_CUDA(m, cudaMemcpyAsync(d_occ, occupied, n, cudaMemcpyHostToDevice, m->stream));
    else ORBM_CUDA(m, cudaMemsetAsync(d_occ, 0, n, m->stream));
  }
  launch_build_grid(F.kps, nullptr, n, 0, 1, min_x, min_y, inv_w, inv_h, d_off, d_items, 0, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  F.occupied = d_occ;
  F.cell_offsets = d_off;
  F.cell_items = d_items;
  F.min_x = min_x;
  F.min_y = min_y;
  F.inv_w = inv_w;
  F.inv_h = inv_h;
  *out = F;
  return ORBX_OK;
}
}  // namespace

int orbm_search_by_projection_map_resident(orbm_matcher* m, const orbx_extractor* ex, int frame, int n,
                                           const float* u_right, const uint8_t* occupied, float min_x, float min_y,
                                           float inv_w, float inv_h, const orbx_mappoints* mps, float th, float nnratio,
                                           int far_points, float th_far, int32_t* assign, int32_t* nmatches) {
  if (!m || !ex || !mps || n < 0 || mps->m < 0 || (n > 0 && !assign)) return mfail(m, ORBX_E_ARG, "bad argument");
  if (nmatches) *nmatches = 0;
  Arena ar(m);
  DevFrame F;
  float sf[kMaxLevels];
  int rc = resident_frame(m, ar, ex, frame, n, u_right, occupied, min_x, min_y, inv_w, inv_h, &F, sf);
  if (rc) return rc;
  return search_map_common(m, ar, F, sf, F.n_levels, u_right != nullptr, mps, th, nnratio, far_points, th_far, n, assign,
                           nmatches);
}

namespace {
int search_frame_common(orbm_matcher* m, Arena& ar, const DevFrame& F, bool has_u_right, const orbx_projected* pts,
                        int max_dist, int check_orientation, int n, int32_t* assign, int32_t* nmatches) {
  const int M = pts->m;
  DevQueries Q{};
  Q.m = M;
  Q.active = nullptr;
  Q.u = ar.upload(pts->u, M);
  Q.v = ar.upload(pts->v, M);
  Q.radius = ar.upload(pts->radius, M);
  Q.min_level = ar.upload(pts->min_level, M);
  Q.max_level = ar.upload(pts->max_level, M);
  Q.u_right = (has_u_right && pts->u_right) ? ar.upload(pts->u_right, M) : (ar.alloc<float>(1), nullptr);
  Q.desc = ar.upload(pts->desc, (size_t)M * 32);
  ResolveArgs R{};
  R.mode = 1;
  R.nnratio = 0.f;
  R.max_dist = max_dist;
  R.check_orientation = check_orientation;
  R.has_obs = ar.upload(pts->has_obs, M);
  R.angle = ar.upload(pts->angle, M);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  return run_search(m, ar, F, Q, R, n, assign, nmatches);
}
}  // namespace

int orbm_search_by_projection_frame(orbm_matcher* m, const orbx_frame_view* f, const orbx_projected* pts,
                                    int max_dist, int check_orientation, int32_t* assign, int32_t* nmatches) {
  if (!m || !f || !pts || f->n < 0 || pts->m < 0 || (f->n > 0 && !assign)) return mfail(m, ORBX_E_ARG, "bad argument");
  if (nmatches) *nmatches = 0;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  const DevFrame F = upload_frame(ar, f);
  return search_frame_common(m, ar, F, f->u_right != nullptr, pts, max_dist, check_orientation, f->n, assign, nmatches);
}

int orbm_search_by_projection_frame_decisions(orbm_matcher* m, const orbx_frame_view* f, const orbx_projected* pts,
                                              int max_dist, int32_t* decisions, int32_t* window_count) {
  if (!m || !f || !pts || f->n < 0 || pts->m < 0 || (pts->m > 0 && !decisions)) return mfail(m, ORBX_E_ARG, "bad argument");
  if (pts->m == 0) return ORBX_OK;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  const DevFrame F = upload_frame(ar, f);
  const int M = pts->m;
  DevQueries Q{};
  Q.m = M;
  Q.active = nullptr;
  Q.u = ar.upload(pts->u, M);
  Q.v = ar.upload(pts->v, M);
  Q.radius = ar.upload(pts->radius, M);
  Q.min_level = ar.upload(pts->min_level, M);
  Q.max_level = ar.upload(pts->max_level, M);
  Q.u_right = (f->u_right && pts->u_right) ? ar.upload(pts->u_right, M) : (ar.alloc<float>(1), nullptr);
  Q.desc = ar.upload(pts->desc, (size_t)M * 32);
  ResolveArgs R{};
  R.mode = 1;
  R.nnratio = 0.f;
  R.max_dist = max_dist;
  R.check_orientation = 0;
  R.has_obs = ar.upload(pts->has_obs, M);
  R.angle = ar.upload(pts->angle, M);
  // |GetFeaturesInArea| per point: the candidate count with nothing occupied and no stereo gate
  int32_t* d_win = nullptr;
  uint8_t* d_zero = nullptr;
  if (window_count) {
    d_win = ar.alloc<int32_t>((size_t)M + 1);
    d_zero = ar.alloc<uint8_t>((size_t)std::max(f->n, 1));
  }
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  if (window_count) {
    DevFrame F0 = F;
    DevQueries Q0 = Q;
    ORBM_CUDA(m, cudaMemsetAsync(d_zero, 0, (size_t)std::max(f->n, 1), m->stream));
    F0.occupied = d_zero;
    F0.u_right = nullptr;
    Q0.u_right = nullptr;
    launch_search_count(F0, Q0, d_win, m->stream);
    ORBM_CUDA(m, cudaMemcpyAsync(window_count, d_win, (size_t)M * 4, cudaMemcpyDeviceToHost, m->stream));
  }
  std::vector<int32_t> assign_unused((size_t)std::max(f->n, 1));
  return run_search(m, ar, F, Q, R, f->n, assign_unused.data(), nullptr, decisions);
}

int orbm_search_by_projection_frame_resident(orbm_matcher* m, const orbx_extractor* ex, int frame, int n,
                                             const float* u_right, const uint8_t* occupied, float min_x, float min_y,
                                             float inv_w, float inv_h, const orbx_projected* pts, int max_dist,
                                             int check_orientation, int32_t* assign, int32_t* nmatches) {
  if (!m || !ex || !pts || n < 0 || pts->m < 0 || (n > 0 && !assign)) return mfail(m, ORBX_E_ARG, "bad argument");
  if (nmatches) *nmatches = 0;
  Arena ar(m);
  DevFrame F;
  float sf[kMaxLevels];
  int rc = resident_frame(m, ar, ex, frame, n, u_right, occupied, min_x, min_y, inv_w, inv_h, &F, sf);
  if (rc) return rc;
  return search_frame_common(m, ar, F, u_right != nullptr, pts, max_dist, check_orientation, n, assign, nmatches);
}

static DevKeyFrame upload_keyframe(Arena& ar, const orbx_keyframe_view* k) {
  DevKeyFrame K{};
  K.n = k->n;
  K.n_levels = k->n_levels;
  K.kps = ar.upload(k->kps, k->n);
  K.desc = ar.upload(k->desc, (size_t)k->n * 32);
  K.u_right = k->u_right ? ar.upload(k->u_right, k->n) : (ar.alloc<float>(1), nullptr);
  K.has_mappoint = ar.upload(k->has_mappoint, k->n);
  K.n_nodes = k->featvec.n_nodes;
  K.node_ids = ar.upload(k->featvec.node_ids, K.n_nodes);
  K.offsets = ar.upload(k->featvec.offsets, (size_t)K.n_nodes + 1);
  K.indices = ar.upload(k->featvec.indices, K.n_nodes > 0 ? (size_t)k->featvec.offsets[K.n_nodes] : 0);
  K.scale_factors = ar.upload(k->scale_factors, k->n_levels);
  K.level_sigma2 = ar.upload(k->level_sigma2, k->n_levels);
  return K;
}

int orbm_set_vocabulary(orbm_matcher* m, const orbx_vocabulary* voc) {
  if (!m || !voc || voc->n_nodes < 1 || voc->depth < 0 || !voc->child_offsets || !voc->descriptors || !voc->word_id ||
      !voc->weight)
    return mfail(m, ORBX_E_ARG, "bad argument");
  const int N = voc->n_nodes;
  const int n_children = voc->child_offsets[N];
  if (voc->child_offsets[0] != 0 || n_children < 0 || (n_children > 0 && !voc->children) ||
      voc->child_offsets[1] <= 0)
    return mfail(m, ORBX_E_ARG, "vocabulary: the root needs children and offsets must start at 0");
  for (int i = 0; i < N; i++)
    if (voc->child_offsets[i + 1] < voc->child_offsets[i]) return mfail(m, ORBX_E_ARG, "vocabulary: offsets must ascend");
  for (int c = 0; c < n_children; c++)
    if (voc->children[c] == 0 || voc->children[c] >= (uint32_t)N)
      return mfail(m, ORBX_E_ARG, "vocabulary: child id out of range");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  const size_t bytes[5] = {(size_t)(N + 1) * 4, (size_t)std::max(n_children, 1) * 4, (size_t)N * 32, (size_t)N * 4,
                           (size_t)N * 8};
  const void* host[5] = {voc->child_offsets, voc->children, voc->descriptors, voc->word_id, voc->weight};
  for (int k = 0; k < 5; k++) {
    cudaError_t e = m->voc_buf[k].reserve(bytes[k]);
    if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(e));
    if (host[k] && (k != 1 || n_children > 0))
      ORBM_CUDA(m, cudaMemcpyAsync(m->voc_buf[k].p, host[k], k == 1 ? (size_t)n_children * 4 : bytes[k],
                                   cudaMemcpyHostToDevice, m->stream));
  }
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  m->voc.n_nodes = N;
  m->voc.depth = voc->depth;
  m->voc.child_offsets = reinterpret_cast<const int32_t*>(m->voc_buf[0].p);
  m->voc.children = reinterpret_cast<const uint32_t*>(m->voc_buf[1].p);
  m->voc.descriptors = reinterpret_cast<const uint8_t*>(m->voc_buf[2].p);
  m->voc.word_id = reinterpret_cast<const uint32_t*>(m->voc_buf[3].p);
  m->voc.weight = reinterpret_cast<const double*>(m->voc_buf[4].p);
  return ORBX_OK;
}

int orbm_bow_transform(orbm_matcher* m, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight,
                       uint32_t* node_id) {
  if (!m || n < 0 || (n > 0 && (!desc || !word_id || !weight || !node_id))) return mfail(m, ORBX_E_ARG, "bad argument");
  if (m->voc.n_nodes < 1) return mfail(m, ORBX_E_ARG, "no vocabulary: call orbm_set_vocabulary first");
  if (n == 0) return ORBX_OK;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  const uint8_t* dd = ar.upload(desc, (size_t)n * 32);
  uint32_t* dw = ar.alloc<uint32_t>(n);
  double* dwt = ar.alloc<double>(n);
  uint32_t* dn = ar.alloc<uint32_t>(n);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_bow_transform(m->voc, dd, n, levelsup, dw, dwt, dn, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  ORBM_CUDA(m, cudaMemcpyAsync(word_id, dw, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaMemcpyAsync(weight, dwt, (size_t)n * 8, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaMemcpyAsync(node_id, dn, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  return ORBX_OK;
}

int orbm_search_for_initialization(orbm_matcher* m, const orbx_frame_view* f1, const orbx_frame_view* f2,
                                   const float* prev_matched_xy, int window_size, float nnratio, int check_orientation,
                                   int32_t* matches12, int32_t* nmatches) {
  if (!m || !f1 || !f2 || f1->n < 0 || f2->n < 0 || (f1->n > 0 && (!matches12 || !prev_matched_xy)))
    return mfail(m, ORBX_E_ARG, "bad argument");
  if (nmatches) *nmatches = 0;
  if (f1->n == 0) return ORBX_OK;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  const int M = f1->n;
  // per-keypoint query set-up = the scalar prologue of the loop body (:637-656): only level-0 keypoints take part, the
  // window is centred on vbPrevMatched with half-size windowSize and restricted to level 0
  std::vector<uint8_t> active(M);
  std::vector<float> u(M), v(M), radius(M, (float)window_size);
  std::vector<int32_t> zero(M, 0);
  for (int i = 0; i < M; i++) {
    active[i] = f1->kps[i].octave <= 0;
    u[i] = prev_matched_xy[2 * i];
    v[i] = prev_matched_xy[2 * i + 1];
  }
  std::vector<uint8_t> free2(std::max(f2->n, 1), 0);
  orbx_frame_view f2v = *f2;
  f2v.occupied = free2.data();  // the static "occupied" test of the other searches does not exist here
  f2v.u_right = nullptr;
  Arena ar(m);
  const DevFrame F2 = upload_frame(ar, &f2v);
  DevQueries Q{};
  Q.m = M;
  Q.active = ar.upload(active.data(), M);
  Q.u = ar.upload(u.data(), M);
  Q.v = ar.upload(v.data(), M);
  Q.radius = ar.upload(radius.data(), M);
  Q.min_level = ar.upload(zero.data(), M);
  Q.max_level = ar.upload(zero.data(), M);
  Q.u_right = nullptr;
  Q.desc = ar.upload(f1->desc, (size_t)M * 32);
  InitArgs A{};
  A.kps1 = ar.upload(f1->kps, M);
  A.nnratio = nnratio;
  A.check_orientation = check_orientation;
  A.n2 = f2->n;
  A.matches12 = ar.alloc<int32_t>(M);
  A.matches21 = ar.alloc<int32_t>(f2->n);
  A.matched_dist = ar.alloc<int32_t>(f2->n);
  A.events = ar.alloc<int32_t>(2 * (size_t)M);
  A.nmatches = ar.alloc<int32_t>(1);
  SearchScratch S{};
  S.counts = ar.alloc<int32_t>((size_t)M + 1);
  S.pre = ar.alloc<int4>(M);
  int32_t* d_total = ar.alloc<int32_t>(1);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  cudaStream_t st = m->stream;
  int32_t total = 0;
  launch_search_count(F2, Q, S.counts, st);
  launch_scan(S.counts, M, d_total, st);
  ORBM_CUDA(m, cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaStreamSynchronize(st));
  S.cand_idx = ar.alloc<int32_t>(total);
  S.cand_dist = ar.alloc<int32_t>(total);
  S.cap_total = total;
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_search_fill(F2, Q, S, st);
  launch_init_resolve(F2, Q, S, A, st);
  ORBM_CUDA(m, cudaGetLastError());
  int32_t nm = 0;
  ORBM_CUDA(m, cudaMemcpyAsync(matches12, A.matches12, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(&nm, A.nmatches, 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaStreamSynchronize(st));
  if (nmatches) *nmatches = nm;
  return ORBX_OK;
}

int orbm_fuse_match(orbm_matcher* m, const orbx_frame_view* kf, const float* inv_level_sigma2,
                    const orbx_projected* pts, int chi2_gate, int32_t* best_idx, int32_t* best_dist) {
  if (!m || !kf || !pts || (chi2_gate && !inv_level_sigma2) || kf->n < 0 || pts->m < 0 ||
      (pts->m > 0 && (!best_idx || !best_dist)))
    return mfail(m, ORBX_E_ARG, "bad argument");
  if (pts->m == 0) return ORBX_OK;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  const int M = pts->m;
  Arena ar(m);
  const DevFrame F = upload_frame(ar, kf);
  DevQueries Q{};
  Q.m = M;
  Q.u = ar.upload(pts->u, M);
  Q.v = ar.upload(pts->v, M);
  Q.radius = ar.upload(pts->radius, M);
  Q.max_level = ar.upload(pts->max_level, M);
  Q.u_right = pts->u_right ? ar.upload(pts->u_right, M) : (ar.alloc<float>(1), nullptr);
  Q.desc = ar.upload(pts->desc, (size_t)M * 32);
  const float* d_inv = ar.upload(inv_level_sigma2, kf->n_levels);
  int32_t* d_bi = ar.alloc<int32_t>(M);
  int32_t* d_bd = ar.alloc<int32_t>(M);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_fuse_match(F, Q, d_inv, chi2_gate, d_bi, d_bd, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  ORBM_CUDA(m, cudaMemcpyAsync(best_idx, d_bi, (size_t)M * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaMemcpyAsync(best_dist, d_bd, (size_t)M * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  return ORBX_OK;
}

namespace {
// every feature of the second view under at most one node (DBoW2's FeatureVector): what makes the nodes independent
bool featvec_is_disjoint(const orbx_keyframe_view* k) {
  const orbx_featvec& v = k->featvec;
  const int total = v.n_nodes > 0 ? v.offsets[v.n_nodes] : 0;
  std::vector<uint8_t> seen(std::max(k->n, 1), 0);
  for (int i = 0; i < total; i++) {
    const uint32_t idx = v.indices[i];
    if (idx >= (uint32_t)k->n || seen[idx]) return false;
    seen[idx] = 1;
  }
  return true;
}

int search_by_bow_common(orbm_matcher* m, const orbx_keyframe_view* kf, const orbx_keyframe_view* second, float nnratio,
                         int check_orientation, int kf_kf, int32_t* matches, int32_t* nmatches, int n_left_f = -1) {
  if (!m || !kf || !second || kf->n < 0 || second->n < 0) return mfail(m, ORBX_E_ARG, "bad argument");
  const int n_out = kf_kf ? kf->n : second->n;
  if (n_out > 0 && !matches) return mfail(m, ORBX_E_ARG, "bad argument");
  if (nmatches) *nmatches = 0;
  if (!featvec_is_disjoint(second)) return mfail(m, ORBX_E_ARG, "FeatureVector lists a feature twice");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  BowArgs A{};
  A.kf = upload_keyframe(ar, kf);
  A.fr = upload_keyframe(ar, second);
  A.nnratio = nnratio;
  A.check_orientation = check_orientation;
  A.kf_kf = kf_kf;
  A.n_left_f = n_left_f;
  A.matches_f = ar.alloc<int32_t>(n_out);
  A.matched2 = ar.alloc<uint8_t>(second->n);
  A.nmatches = ar.alloc<int32_t>(1);
  A.node_match = ar.alloc<int32_t>(A.kf.n_nodes);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_search_by_bow(A, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  int32_t nm = 0;
  if (n_out > 0)
    ORBM_CUDA(m, cudaMemcpyAsync(matches, A.matches_f, (size_t)n_out * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaMemcpyAsync(&nm, A.nmatches, 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  if (nmatches) *nmatches = nm;
  return ORBX_OK;
}
}  // namespace

int orbm_search_by_bow(orbm_matcher* m, const orbx_keyframe_view* kf, const orbx_keyframe_view* frame, float nnratio,
                       int check_orientation, int32_t* matches_f, int32_t* nmatches) {
  return search_by_bow_common(m, kf, frame, nnratio, check_orientation, 0, matches_f, nmatches);
}

int orbm_search_by_bow_fisheye(orbm_matcher* m, const orbx_keyframe_view* kf, const orbx_keyframe_view* frame,
                               int n_left_frame, float nnratio, int check_orientation, int32_t* matches_f,
                               int32_t* nmatches) {
  if (!frame || n_left_frame < 0 || n_left_frame > frame->n) return mfail(m, ORBX_E_ARG, "bad argument");
  return search_by_bow_common(m, kf, frame, nnratio, check_orientation, 0, matches_f, nmatches, n_left_frame);
}

int orbm_search_by_bow_kf(orbm_matcher* m, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, float nnratio,
                          int check_orientation, int32_t* matches12, int32_t* nmatches) {
  return search_by_bow_common(m, kf1, kf2, nnratio, check_orientation, 1, matches12, nmatches);
}

int orbm_search_for_triangulation(orbm_matcher* m, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2,
                                  const float* F12, float ep_x, float ep_y, int only_stereo, int coarse,
                                  int check_orientation, int32_t* matches12, int32_t* nmatches) {
  if (!m || !kf1 || !kf2 || !F12 || kf1->n < 0 || kf2->n < 0 || (kf1->n > 0 && !matches12))
    return mfail(m, ORBX_E_ARG, "bad argument");
  if (nmatches) *nmatches = 0;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  TriArgs A{};
  A.k1 = upload_keyframe(ar, kf1);
  A.k2 = upload_keyframe(ar, kf2);
  memcpy(A.F12, F12, sizeof(A.F12));
  A.ep_x = ep_x;
  A.ep_y = ep_y;
  A.only_stereo = only_stereo;
  A.coarse = coarse;
  A.check_orientation = check_orientation;
  A.matches12 = ar.alloc<int32_t>(kf1->n);
  A.nmatches = ar.alloc<int32_t>(1);
  A.node_match = ar.alloc<int32_t>(A.k1.n_nodes);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_triangulation(A, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  int32_t nm = 0;
  if (kf1->n > 0)
    ORBM_CUDA(m, cudaMemcpyAsync(matches12, A.matches12, (size_t)kf1->n * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaMemcpyAsync(&nm, A.nmatches, 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  if (nmatches) *nmatches = nm;
  return ORBX_OK;
}

int orbm_triangulation_candidates(orbm_matcher* m, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2,
                                  int32_t* offsets, int32_t* cand_idx2, int32_t* cand_dist, int32_t cap, int32_t* total) {
  if (!m || !kf1 || !kf2 || kf1->n < 0 || kf2->n < 0 || !offsets || !total || cap < 0 || (cap > 0 && (!cand_idx2 || !cand_dist)))
    return mfail(m, ORBX_E_ARG, "bad argument");
  *total = 0;
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  TriArgs A{};
  A.k1 = upload_keyframe(ar, kf1);
  A.k2 = upload_keyframe(ar, kf2);
  A.node_match = ar.alloc<int32_t>(A.k1.n_nodes);
  int32_t* d_off = ar.alloc<int32_t>(kf1->n + 1);
  int32_t* d_total = ar.alloc<int32_t>(1);
  int32_t* d_idx = ar.alloc<int32_t>(cap);
  int32_t* d_dist = ar.alloc<int32_t>(cap);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  launch_triangulation_candidates(A, d_off, d_total, nullptr, nullptr, 0, false, m->stream);
  launch_triangulation_candidates(A, d_off, d_total, d_idx, d_dist, cap, true, m->stream);
  ORBM_CUDA(m, cudaGetLastError());
  int32_t t = 0;
  ORBM_CUDA(m, cudaMemcpyAsync(offsets, d_off, (size_t)(kf1->n + 1) * 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaMemcpyAsync(&t, d_total, 4, cudaMemcpyDeviceToHost, m->stream));
  ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  *total = t;
  const int nw = t < cap ? t : cap;
  if (nw > 0) {
    ORBM_CUDA(m, cudaMemcpyAsync(cand_idx2, d_idx, (size_t)nw * 4, cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaMemcpyAsync(cand_dist, d_dist, (size_t)nw * 4, cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
  }
  return t > cap ? mfail(m, ORBX_E_CAPACITY, "candidate buffer too small: see *total") : ORBX_OK;
}

}  // extern "C"
