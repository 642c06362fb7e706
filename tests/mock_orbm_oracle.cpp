// TEST-ONLY stand-in of the orbm C ABI entry points the reference-side shim bodies (shim/*.cc) call, answering with the
// CPU oracle instead of the CUDA library. It exists so that the shim's flatten / scatter glue — the code that turns a
// Frame / KeyFrame / MapPoint graph into the flat views and writes the answers back — can run HERE, on stand-in frames,
// against the reference's own ORBmatcher.cc (tests/test_shim_bodies_vs_reference_source.py). Never linked into the product.
#include <cstring>

#include "../include/orbm.h"
#include "../include/orbx.h"
#include "../oracle/orbref.h"

#include <vector>

// the mock context keeps its own copy of the vocabulary, as the library keeps one on the device
struct orbm_matcher {
  std::vector<int32_t> child_offsets;
  std::vector<uint32_t> children, word_id;
  std::vector<uint8_t> descriptors;
  std::vector<double> weight;
  orbx_vocabulary voc{};
};
// the extractor handle of the mock: the oracle's extractor, whose last call holds the pyramid the stereo matcher reads
struct orbx_extractor {
  orbref_extractor* r;
  int nfeatures, nlevels;
};

extern "C" {
int orbm_create(orbm_matcher** out, int) {
  *out = new orbm_matcher();
  return ORBX_OK;
}
void orbm_destroy(orbm_matcher* m) { delete m; }
const char* orbm_last_error(const orbm_matcher*) { return "mock"; }

int orbm_search_by_projection_map(orbm_matcher*, const orbx_frame_view* f, const orbx_mappoints* mps, float th,
                                  float nnratio, int far_points, float th_far, int32_t* assign, int32_t* nmatches) {
  const int n = orbref_search_by_projection_map(f, mps, th, nnratio, far_points, th_far, assign);
  if (nmatches) *nmatches = n;
  return ORBX_OK;
}
int orbm_set_vocabulary(orbm_matcher* m, const orbx_vocabulary* v) {
  const int N = v->n_nodes, nc = v->child_offsets[N];
  m->child_offsets.assign(v->child_offsets, v->child_offsets + N + 1);
  m->children.assign(v->children, v->children + nc);
  m->word_id.assign(v->word_id, v->word_id + N);
  m->descriptors.assign(v->descriptors, v->descriptors + (size_t)N * 32);
  m->weight.assign(v->weight, v->weight + N);
  m->voc = orbx_vocabulary{N, v->depth, m->child_offsets.data(), m->children.data(), m->descriptors.data(),
                           m->word_id.data(), m->weight.data()};
  return ORBX_OK;
}
int orbm_bow_transform(orbm_matcher* m, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight,
                       uint32_t* node_id) {
  if (m->voc.n_nodes < 1) return ORBX_E_ARG;
  orbref_bow_transform(&m->voc, desc, n, levelsup, word_id, weight, node_id);
  return ORBX_OK;
}
int orbm_is_in_frustum(orbm_matcher*, const orbx_frustum* fr, const orbx_local_map* map, int map_index,
                       float viewing_cos_limit, uint8_t* track_in_view, float* proj_x, float* proj_y, float* proj_xr,
                       int32_t* level, float* view_cos, float* depth, int32_t* n_in_view) {
  orbref_is_in_frustum(fr, map, map_index, viewing_cos_limit, track_in_view, proj_x, proj_y, proj_xr, level, view_cos, depth);
  int n = 0;
  for (int i = 0; i < map->m; i++) n += track_in_view[i] != 0;
  if (n_in_view) *n_in_view = n;
  return ORBX_OK;
}
int orbm_search_by_projection_map_fisheye(orbm_matcher*, const orbx_fisheye_view* f, const orbx_mappoints* mps,
                                          const orbx_mappoints_right* mr, float th, float nnratio, int far_points,
                                          float th_far, int32_t* assign, int32_t* nmatches) {
  const int n = orbref_search_by_projection_map_fisheye(f, mps, mr, th, nnratio, far_points, th_far, assign);
  if (nmatches) *nmatches = n;
  return ORBX_OK;
}
int orbm_search_by_projection_frame_decisions(orbm_matcher*, const orbx_frame_view* f, const orbx_projected* pts,
                                              int max_dist, int32_t* decisions, int32_t* window_count) {
  orbref_search_by_projection_frame_decisions(f, pts, max_dist, decisions, window_count);
  return 0;
}
int orbm_search_by_projection_frame(orbm_matcher*, const orbx_frame_view* f, const orbx_projected* pts, int max_dist,
                                    int check_orientation, int32_t* assign, int32_t* nmatches) {
  const int n = orbref_search_by_projection_frame(f, pts, max_dist, check_orientation, assign);
  if (nmatches) *nmatches = n;
  return ORBX_OK;
}
int orbm_search_for_triangulation(orbm_matcher*, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2,
                                  const float* F12, float ep_x, float ep_y, int only_stereo, int coarse,
                                  int check_orientation, int32_t* matches12, int32_t* nmatches) {
  const int n = orbref_search_for_triangulation(kf1, kf2, F12, ep_x, ep_y, only_stereo, coarse, check_orientation, matches12);
  if (nmatches) *nmatches = n;
  return ORBX_OK;
}
int orbm_triangulation_candidates(orbm_matcher*, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2,
                                  int32_t* offsets, int32_t* cand_idx2, int32_t* cand_dist, int32_t cap, int32_t* total) {
  *total = orbref_triangulation_candidates(kf1, kf2, offsets, cand_idx2, cand_dist, cap);
  return *total > cap ? ORBX_E_CAPACITY : ORBX_OK;
}
int orbm_search_by_bow(orbm_matcher*, const orbx_keyframe_view* kf, const orbx_keyframe_view* frame, float nnratio,
                       int check_orientation, int32_t* matches_f, int32_t* nmatches) {
  const int n = orbref_search_by_bow(kf, frame, nnratio, check_orientation, matches_f);
  if (nmatches) *nmatches = n;
  return ORBX_OK;
}
int orbm_search_by_bow_fisheye(orbm_matcher*, const orbx_keyframe_view* kf, const orbx_keyframe_view* frame,
                               int n_left_frame, float nnratio, int check_orientation, int32_t* matches_f,
                               int32_t* nmatches) {
  const int n = orbref_search_by_bow_fisheye(kf, frame, n_left_frame, nnratio, check_orientation, matches_f);
  if (nmatches) *nmatches = n;
  return ORBX_OK;
}
int orbm_search_by_bow_kf(orbm_matcher*, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, float nnratio,
                          int check_orientation, int32_t* matches12, int32_t* nmatches) {
  const int n = orbref_search_by_bow_kf(kf1, kf2, nnratio, check_orientation, matches12);
  if (nmatches) *nmatches = n;
  return ORBX_OK;
}
int orbm_fuse_match(orbm_matcher*, const orbx_frame_view* kf, const float* inv_level_sigma2, const orbx_projected* pts,
                    int chi2_gate, int32_t* best_idx, int32_t* best_dist) {
  orbref_fuse_match(kf, inv_level_sigma2, pts, chi2_gate, best_idx, best_dist);
  return ORBX_OK;
}
int orbm_search_for_initialization(orbm_matcher*, const orbx_frame_view* f1, const orbx_frame_view* f2,
                                   const float* prev_matched_xy, int window_size, float nnratio, int check_orientation,
                                   int32_t* matches12, int32_t* nmatches) {
  const int n = orbref_search_for_initialization(f1, f2, prev_matched_xy, window_size, nnratio, check_orientation, matches12);
  if (nmatches) *nmatches = n;
  return ORBX_OK;
}
int orbm_assign_features_to_grid(orbm_matcher*, const orbx_kp* kps, int n, float min_x, float min_y, float inv_w,
                                 float inv_h, int32_t* cell_offsets, int32_t* cell_items) {
  orbref_build_grid(kps, n, min_x, min_y, inv_w, inv_h, cell_offsets, cell_items);
  return ORBX_OK;
}
int orbm_distinctive_descriptors(orbm_matcher*, const uint8_t* desc, const int32_t* offsets, int n_points,
                                 int32_t* best_idx) {
  for (int i = 0; i < n_points; i++)
    best_idx[i] = orbref_distinctive_descriptor(desc + (size_t)offsets[i] * 32, offsets[i + 1] - offsets[i]);
  return ORBX_OK;
}

// ---- the part of include/orbx.h that shim/ORBextractor.cc calls, and orbm_stereo_match for shim/FrameStereo_orbx.cc ----
int orbx_extractor_create(orbx_extractor** out, int, int nfeatures, float scale_factor, int nlevels, int ini_th_fast,
                          int min_th_fast, int) {
  *out = new orbx_extractor{orbref_extractor_create(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast), nfeatures,
                            nlevels};
  return ORBX_OK;
}
void orbx_extractor_destroy(orbx_extractor* ex) {
  if (!ex) return;
  orbref_extractor_destroy(ex->r);
  delete ex;
}
const char* orbx_last_error(const orbx_extractor*) { return "mock"; }
int orbx_extractor_levels(const orbx_extractor* ex) { return ex->nlevels; }
int orbx_extractor_tables(const orbx_extractor* ex, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int32_t* features_per_level) {
  orbref_extractor_tables(ex->r, scale, inv_scale, sigma2, inv_sigma2, features_per_level, nullptr);
  return ORBX_OK;
}
int orbx_extractor_capacity(const orbx_extractor* ex) { return ex->nfeatures + 16 * ex->nlevels; }
int orbx_extract(orbx_extractor* ex, const uint8_t* image, int width, int height, int stride, int lap0, int lap1,
                 orbx_kp* kps, uint8_t* desc, int cap, int32_t* n_out, int32_t* mono_index) {
  int n = 0, mono = 0;
  const int rc = orbref_extract(ex->r, image, width, height, stride, lap0, lap1, kps, desc, cap, &n, &mono);
  *n_out = n;
  *mono_index = mono;
  return rc == 0 ? ORBX_OK : (rc == -1 ? ORBX_E_EMPTY : ORBX_E_CAPACITY);
}
int orbx_level_size(const orbx_extractor* ex, int level, int* width, int* height) {
  orbref_level_dims(ex->r, level, width, height);
  return ORBX_OK;
}
int orbx_download_pyramid(orbx_extractor* ex, int, int level, uint8_t* dst, int dst_stride) {
  int w = 0, h = 0, st = 0;
  orbref_level_dims(ex->r, level, &w, &h);
  const uint8_t* src = orbref_level_bordered(ex->r, level, &st);
  for (int y = 0; y < h + 38; y++) memcpy(dst + (size_t)y * dst_stride, src + (size_t)y * st, (size_t)w + 38);
  return ORBX_OK;
}
int orbm_knn2(orbm_matcher*, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx1, int32_t* d1, int32_t* idx2,
              int32_t* d2) {
  orbref_knn2(q, nq, t, nt, idx1, d1, idx2, d2);
  return ORBX_OK;
}
int orbm_stereo_match(orbm_matcher*, const orbx_extractor* left, const orbx_extractor* right, int, const orbx_kp* kps_l,
                      const uint8_t* desc_l, int n_l, const orbx_kp* kps_r, const uint8_t* desc_r, int n_r, float mbf,
                      float mb, float* u_right, float* depth, int32_t* n_matched) {
  const int n = orbref_stereo_match(left->r, right->r, kps_l, desc_l, n_l, kps_r, desc_r, n_r, mbf, mb, u_right, depth);
  if (n_matched) *n_matched = n;
  return ORBX_OK;
}
}
