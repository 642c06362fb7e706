// orbm_track_api.cu — Frame::isInFrustum and the batched Tracking::SearchLocalPoints entry points of include/orbm.h
// (kernels: k_track.cu). No result is computed on the host.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "orbm_handle.h"

using namespace orbx;
static_assert(sizeof(orbx_frustum) == 104, "orbx_frustum is 26 words");

namespace orbm_detail {

// Scratch of one batched tracking search in m->track_buf[slot]; fills the scratch pointers of A (n_frames, cap, m set).
int track_prepare(orbm_matcher* m, int slot, TrackArgs* A) {
  const size_t F = (size_t)A->n_frames, M = (size_t)std::max(A->m, 1);
  const size_t bytes[kTrackBufs] = {F * kTrackOff16 * 2,       F * A->cap * 16, F * M * sizeof(float4), F * M * 4,
                                    F * M * sizeof(int2),    F * M * sizeof(int4), F * (size_t)A->cand_cap * 4,
                                    F * 4,                   F * M * 4,      0, 0, 0};
  DevBuf* b = m->track_buf[slot];
  for (int k = 0; k < kTrackBufs; k++) {
    if (!bytes[k]) continue;
    cudaError_t e = b[k].reserve(bytes[k]);
    if (e != cudaSuccess) return mfail(m, ORBX_E_CUDA, std::string("track scratch: ") + cudaGetErrorString(e));
  }
  A->grid_off16 = static_cast<uint16_t*>(b[0].p);
  A->grid_rec = static_cast<uint4*>(b[1].p);
  A->q = static_cast<float4*>(b[2].p);
  A->q_level = static_cast<int32_t*>(b[3].p);
  A->seg = static_cast<int2*>(b[4].p);
  A->pre = static_cast<int4*>(b[5].p);
  A->cand = static_cast<uint32_t*>(b[6].p);
  A->cand_total = static_cast<int32_t*>(b[7].p);
  A->dec = static_cast<int32_t*>(b[8].p);
  return ORBX_OK;
}

int track_fill_params(orbm_matcher* m, const orbx_extractor* ex, const orbx_local_map* maps,
                      const orbx_track_params* prm, int cap, TrackArgs* A) {
  if (!ex || !maps || !prm) return mfail(m, ORBX_E_ARG, "bad argument");
  if (ex->device != m->device) return mfail(m, ORBX_E_ARG, "extractor and matcher must live on the same device");
  if (cap < 1 || cap >= 65536) return mfail(m, ORBX_E_ARG, "cap must be in [1, 65535]");
  if (maps->m < 0 || maps->n_maps < 1) return mfail(m, ORBX_E_ARG, "bad local map");
  if (maps->m > 0 && (!maps->pos || !maps->normal || !maps->min_dist || !maps->max_dist || !maps->has_obs || !maps->desc))
    return mfail(m, ORBX_E_ARG, "local map array missing");
  if (ex->nlevels > kMaxLevels) return mfail(m, ORBX_E_ARG, "too many levels");
  A->cap = cap;
  A->n_levels = ex->nlevels;
  float s2[kMaxLevels];
  int quota[kMaxLevels];
  make_tables(ex->nfeatures, ex->scale_factor, ex->nlevels, A->scale_factors, s2, quota);  // mvScaleFactors
  A->m = maps->m;
  A->n_maps = maps->n_maps;
  A->viewing_cos_limit = prm->viewing_cos_limit;
  A->th = prm->th;
  A->nnratio = prm->nnratio;
  A->far_points = prm->far_points;
  A->th_far = prm->th_far;
  A->min_x = prm->min_x;
  A->min_y = prm->min_y;
  A->inv_w = prm->inv_w;
  A->inv_h = prm->inv_h;
  // Default slab: 16 records per point at th <= 2 (a th = 1 window holds 1-6 records), growing with the window area
  // (th^2), never more than every keypoint for every point
  long long cc = prm->cand_per_frame;
  if (cc <= 0) {
    const double area = std::max(1.0, 0.25 * (double)prm->th * (double)prm->th);
    cc = std::min(std::max((long long)(16.0 * area * maps->m), 65536LL), std::max((long long)maps->m * cap, 65536LL));
  }
  if (cc > (1LL << 30)) return mfail(m, ORBX_E_ARG, "cand_per_frame too large");
  A->cand_cap = (int)cc;
  return ORBX_OK;
}

void track_set_map(const orbx_local_map* d, TrackArgs* A) {
  A->pos = d->pos;
  A->normal = d->normal;
  A->min_dist = d->min_dist;
  A->max_dist = d->max_dist;
  A->skip = d->skip;
  A->has_obs = d->has_obs;
  A->mdesc = d->desc;
}

}  // namespace orbm_detail

extern "C" {

int orbm_is_in_frustum(orbm_matcher* m, const orbx_frustum* fr, const orbx_local_map* map, int map_index,
                       float viewing_cos_limit, uint8_t* track_in_view, float* proj_x, float* proj_y, float* proj_xr,
                       int32_t* level, float* view_cos, float* depth, int32_t* n_in_view) {
  if (!m || !fr || !map || map->m < 0 || map->n_maps < 1 || map_index < 0 || map_index >= map->n_maps)
    return mfail(m, ORBX_E_ARG, "bad argument");
  const int M = map->m;
  if (n_in_view) *n_in_view = 0;
  if (M == 0) return ORBX_OK;
  if (!map->pos || !map->normal || !map->min_dist || !map->max_dist || !track_in_view || !proj_x || !proj_y ||
      !proj_xr || !level || !view_cos || !depth)
    return mfail(m, ORBX_E_ARG, "bad argument");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  Arena ar(m);
  const size_t base = (size_t)map_index * M;
  TrackArgs A{};
  A.n_frames = 1;
  A.m = M;
  A.n_maps = 1;
  A.pos = ar.upload(map->pos + 3 * base, (size_t)M * 3);
  A.normal = ar.upload(map->normal + 3 * base, (size_t)M * 3);
  A.min_dist = ar.upload(map->min_dist + base, M);
  A.max_dist = ar.upload(map->max_dist + base, M);
  A.skip = map->skip ? ar.upload(map->skip + base, M) : nullptr;
  A.frustums = ar.upload(fr, 1);
  A.viewing_cos_limit = viewing_cos_limit;
  A.th = 1.f;
  A.n_levels = 0;  // no query is produced for a search: only the projection outputs are wanted
  A.o_in_view = ar.alloc<uint8_t>(M);
  A.o_proj_x = ar.alloc<float>(M);
  A.o_proj_y = ar.alloc<float>(M);
  A.o_proj_xr = ar.alloc<float>(M);
  A.o_view_cos = ar.alloc<float>(M);
  A.o_depth = ar.alloc<float>(M);
  A.o_level = ar.alloc<int32_t>(M);
  A.q = ar.alloc<float4>(M);
  A.q_level = ar.alloc<int32_t>(M);
  A.n_in_view = ar.alloc<int32_t>(1);
  if (ar.sync_uploads() != cudaSuccess) return mfail(m, ORBX_E_CUDA, cudaGetErrorString(ar.err));
  cudaStream_t st = m->stream;
  // outputs the reference leaves untouched keep the caller's bytes: they make the round trip
  ORBM_CUDA(m, cudaMemcpyAsync(A.o_proj_xr, proj_xr, (size_t)M * 4, cudaMemcpyHostToDevice, st));
  ORBM_CUDA(m, cudaMemcpyAsync(A.o_view_cos, view_cos, (size_t)M * 4, cudaMemcpyHostToDevice, st));
  ORBM_CUDA(m, cudaMemcpyAsync(A.o_depth, depth, (size_t)M * 4, cudaMemcpyHostToDevice, st));
  ORBM_CUDA(m, cudaMemcpyAsync(A.o_level, level, (size_t)M * 4, cudaMemcpyHostToDevice, st));
  ORBM_CUDA(m, cudaMemcpyAsync(A.o_proj_x, proj_x, (size_t)M * 4, cudaMemcpyHostToDevice, st));
  ORBM_CUDA(m, cudaMemcpyAsync(A.o_proj_y, proj_y, (size_t)M * 4, cudaMemcpyHostToDevice, st));
  launch_frustum_batch(A, st);
  ORBM_CUDA(m, cudaGetLastError());
  int32_t nv = 0;
  ORBM_CUDA(m, cudaMemcpyAsync(track_in_view, A.o_in_view, (size_t)M, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(proj_x, A.o_proj_x, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(proj_y, A.o_proj_y, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(proj_xr, A.o_proj_xr, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(view_cos, A.o_view_cos, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(depth, A.o_depth, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(level, A.o_level, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaMemcpyAsync(&nv, A.n_in_view, 4, cudaMemcpyDeviceToHost, st));
  ORBM_CUDA(m, cudaStreamSynchronize(st));
  if (n_in_view) *n_in_view = nv;
  return ORBX_OK;
}

int orbm_track_local_map_batch_device(orbm_matcher* m, const orbx_extractor* ex, int n_frames, const orbx_kp* d_kps,
                                      const uint8_t* d_desc, const int32_t* d_n, int cap, const float* d_u_right,
                                      const uint8_t* d_occupied, const orbx_frustum* d_frustums,
                                      const orbx_local_map* maps, const int32_t* d_map_index,
                                      const orbx_track_params* prm, int32_t* d_assign, int32_t* d_nmatches,
                                      int32_t* d_n_in_view, int32_t* d_status, void* cuda_stream) {
  if (!m || n_frames < 1 || !d_kps || !d_desc || !d_n || !d_frustums || !d_assign || !d_nmatches || !d_n_in_view ||
      !d_status)
    return mfail(m, ORBX_E_ARG, "bad argument");
  ORBM_CUDA(m, cudaSetDevice(m->device));
  TrackArgs A{};
  int rc = track_fill_params(m, ex, maps, prm, cap, &A);
  if (rc) return rc;
  A.n_frames = n_frames;
  A.kps = d_kps;
  A.desc = d_desc;
  A.n = d_n;
  A.u_right = d_u_right;
  A.occupied = d_occupied;
  A.frustums = d_frustums;
  A.map_index = d_map_index;
  track_set_map(maps, &A);
  if ((rc = track_prepare(m, kLanes, &A)) != 0) return rc;
  A.assign = d_assign;
  A.nmatches = d_nmatches;
  A.n_in_view = d_n_in_view;
  A.status = d_status;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : m->stream;
  launch_frustum_batch(A, st);
  launch_track_search(A, st);
  ORBM_CUDA(m, cudaGetLastError());
  return ORBX_OK;
}

}  // extern "C"
