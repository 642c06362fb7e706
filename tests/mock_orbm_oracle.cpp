// TEST-ONLY stand-in of the orbm C ABI entry points the reference-side shim bodies (shim/*.cc) call, answering with the
// CPU oracle instead of the CUDA library. It exists so that the shim's flatten / scatter glue — the code that turns a
// Frame / KeyFrame / MapPoint graph into the flat views and writes the answers back — can run HERE, on stand-in frames,
// against the reference's own ORBmatcher.cc (tests/test_shim_bodies_vs_reference_source.py). Never linked into the product.
#include <cstring>

#include "../include/orbm.h"
#include "../oracle/orbref.h"

struct orbm_matcher { int unused; };

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
int orbm_search_by_bow(orbm_matcher*, const orbx_keyframe_view* kf, const orbx_keyframe_view* frame, float nnratio,
                       int check_orientation, int32_t* matches_f, int32_t* nmatches) {
  const int n = orbref_search_by_bow(kf, frame, nnratio, check_orientation, matches_f);
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
}
