// orbx_match.cuh — internal launch interface of the matcher kernels (k_match.cu, k_search.cu) used by orbm_api.cu.
#ifndef ORBX_MATCH_CUH_
#define ORBX_MATCH_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orbx_types.h"
#include "orbx_math.h"

namespace orbx {

constexpr int ORBM_TH_HIGH_I = 100;  // ORBmatcher::TH_HIGH   src/ORBmatcher.cc:35
constexpr int ORBM_TH_LOW_I = 50;    // ORBmatcher::TH_LOW    src/ORBmatcher.cc:36
constexpr int kHistoLength = 30;     // HISTO_LENGTH          src/ORBmatcher.cc:37

// raw pyramid of one extractor's last call (border-less levels)
struct PyrView {
  const uint8_t* base[kMaxLevels];
  int64_t fstride[kMaxLevels];
  int pitch[kMaxLevels];
  int w[kMaxLevels];
  int h[kMaxLevels];
};

struct StereoArgs {
  PyrView left, right;
  float scale[kMaxLevels], inv_scale[kMaxLevels];  // Frame::mvScaleFactors / mvInvScaleFactors (left extractor)
  int nlevels;
  int frame0;                     // pair p reads frame frame0 + p of both pyramids
  const orbx_kp *kps_l, *kps_r;   // [pairs][cap]
  const uint8_t *desc_l, *desc_r; // [pairs][cap][32]
  const int32_t *n_l, *n_r;       // [pairs] device counts, or NULL -> n_l_host / n_r_host
  int n_l_host, n_r_host;
  int cap;
  float mbf, mb;
  float *u_right, *depth;         // [pairs][cap]
  int32_t* sad;                   // [pairs][cap] scratch: accepted SAD or -1
  int32_t* n_matched;             // [pairs]
};

int knn2_splits(int nq, int nt);
void launch_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int4* partial, int splits, int32_t* idx1,
                 int32_t* d1, int32_t* idx2, int32_t* d2, cudaStream_t st);
void launch_knn2_merge(const int4* partial, int nq, int splits, int32_t* idx1, int32_t* d1, int32_t* idx2,
                       int32_t* d2, cudaStream_t st);
// k_knn2_tc.cu: the tcgen05 (s8 GEMM) form for large sets; partial = [splits][nq] like launch_knn2's
bool knn2_tc_eligible(int nq, int nt);
size_t knn2_tc_expanded_bytes(int rows);
bool knn2_tc_expands_queries();  // false: the query tile is built in tensor memory from the raw descriptors
int knn2_tc_splits(int nq, int nt, int* tiles_per_split);
cudaError_t launch_knn2_tc(const uint8_t* q, int nq, const uint8_t* t, int nt, int8_t* expanded_q, int8_t* expanded_t,
                           int4* partial, int splits, int tiles_per_split, cudaStream_t st);
void launch_desc_dist(const uint8_t* a, const uint8_t* b, int n, int32_t* out, cudaStream_t st);
// TemplatedVocabulary::transform per feature; the vocabulary arrays are device resident (orbm_vocabulary handle)
struct DevVocabulary {
  int n_nodes, depth;
  const int32_t* child_offsets;
  const uint32_t* children;
  const uint8_t* descriptors;
  const uint32_t* word_id;
  const double* weight;
};
void launch_bow_transform(const DevVocabulary& V, const uint8_t* desc, int n, int levelsup, uint32_t* word_id,
                          double* weight, uint32_t* node_id, cudaStream_t st);
// MapPoint::ComputeDistinctiveDescriptors for n_points CSR lists of descriptors
void launch_distinctive(const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best, cudaStream_t st);
void launch_stereo(const StereoArgs& A, int n_pairs, int max_rows, cudaStream_t st);

// ---- guided searches (k_search.cu) ----
// Device copies of the orbx_frame_view / orbx_mappoints / orbx_projected / orbx_keyframe_view arrays.
struct DevFrame {
  int n, n_levels;
  const orbx_kp* kps;
  const uint8_t* desc;
  const float* u_right;      // or NULL
  const uint8_t* occupied;
  const int32_t* cell_offsets;
  const int32_t* cell_items;
  float min_x, min_y, inv_w, inv_h;
  const float* scale_factors;
};

// One search query per projected point: centre, window half-size, level window, optional stereo gate, descriptor.
struct DevQueries {
  int m;
  const uint8_t* active;     // or NULL = all
  const float *u, *v, *radius;
  const int32_t *min_level, *max_level;
  const float* u_right;      // or NULL
  const uint8_t* desc;
};

struct SearchScratch {
  int32_t* counts;    // [m + 1] candidates per query, then exclusive offsets
  int32_t* cand_idx;  // [total]
  int32_t* cand_dist; // [total] dist | octave << 16
  int4* pre;          // [m] unconstrained best / second best: (d1, pos1, d2, pos2), pos = -1 when absent
  int64_t cap_total;
};

void launch_search_count(const DevFrame& F, const DevQueries& Q, int32_t* counts, cudaStream_t st);
void launch_scan(int32_t* counts, int m, int32_t* total_out, cudaStream_t st);
void launch_search_fill(const DevFrame& F, const DevQueries& Q, const SearchScratch& S, cudaStream_t st);
// mode 0: SearchByProjection(Frame&, vector<MapPoint*>) — ratio test with levels; mode 1: best only + rotation check
struct ResolveArgs {
  int mode;
  float nnratio;
  int max_dist;
  int check_orientation;
  const uint8_t* has_obs;   // [m]
  const float* angle;       // [m] (mode 1)
  int32_t* assign;          // [n]
  int32_t* nmatches;        // [1]
  uint8_t* occ;             // [n] scratch, initialised from DevFrame::occupied
  int32_t* events;          // [2 * m] scratch (mode 1): accepted (idx, bin)
  int32_t* dec;             // [m] scratch: the keypoint each point takes (-1 = none), iterated to the fixed point
};
void launch_search_resolve(const DevFrame& F, const DevQueries& Q, const SearchScratch& S, const ResolveArgs& R,
                           cudaStream_t st);

// ---- batched local-map tracking search (k_track.cu): Tracking::SearchLocalPoints for the frames of one extract batch
// Frame f: keypoints kps + f * cap (n[f] of them), descriptors desc + f * cap * 32, its pose frustums[f], its local map
// map_index[f] (NULL: (map_f0 + f) % n_maps). Everything is device memory; nothing is synchronised.
constexpr int kTrackOff16 = ORBX_GRID_COLS * ORBX_GRID_ROWS + 4;  // u16 cell offsets per frame (3073 used)
struct TrackArgs {
  int n_frames, cap;
  const orbx_kp* kps;
  const uint8_t* desc;
  const int32_t* n;
  const float* u_right;     // [F][cap] mvuRight, or NULL (monocular)
  const uint8_t* occupied;  // [F][cap] or NULL = no keypoint holds a MapPoint yet
  float min_x, min_y, inv_w, inv_h;  // grid geometry (Frame::mnMinX / mnMinY / mfGridElement*Inv)
  float scale_factors[kMaxLevels];
  int n_levels;
  // local maps: [n_maps][m]
  int m, n_maps;
  const float *pos, *normal, *min_dist, *max_dist;
  const uint8_t *skip, *has_obs, *mdesc;
  const int32_t* map_index;
  int map_f0;  // map_index == NULL: frame f uses map (map_f0 + f) % n_maps
  const orbx_frustum* frustums;  // [F]
  float viewing_cos_limit, th, nnratio, th_far;
  int far_points;
  // optional isInFrustum outputs, [F][m] each (any may be NULL)
  uint8_t* o_in_view;
  float *o_proj_x, *o_proj_y, *o_proj_xr, *o_view_cos, *o_depth;
  int32_t* o_level;
  // scratch
  uint16_t* grid_off16;   // [F][kTrackOff16] cell offsets (cap < 65536)
  uint4* grid_rec;        // [F][cap] per keypoint IN CELL ORDER: x, y, u_right (-1 = none), keypoint | octave << 16 | occupied << 31
  float4* q;              // [F][m] projected query: (u, v, u_right, radius)
  int32_t* q_level;       // [F][m] predicted level, -1 = not searched
  int2* seg;              // [F][m] (first candidate, count) inside the frame's candidate slab
  int4* pre;              // [F][m] unconstrained best / second best: (d1, pos1, d2, pos2), pos = -1 when absent
  uint32_t* cand;         // [F][cand_cap] dist << 20 | octave << 16 | keypoint
  int cand_cap;
  int32_t* cand_total;    // [F] (zeroed by the launcher)
  int32_t* dec;           // [F][m]
  // results
  int32_t* assign;        // [F][cap] index of the MapPoint written to mvpMapPoints[i], or -1
  int32_t* nmatches;      // [F] SearchByProjection's return value
  int32_t* n_in_view;     // [F] nToMatch (points isInFrustum accepted)
  int32_t* status;        // [F] 0, or ORBX_E_CAPACITY when the candidate slab overflowed
};
void launch_frustum_batch(const TrackArgs& A, cudaStream_t st);
void launch_track_search(const TrackArgs& A, cudaStream_t st);  // grid -> enumerate -> resolve
size_t track_resolve_smem(int cap);

// SearchByProjection(Frame&, vector<MapPoint*>) on a two-camera Frame (Nleft != -1, src/ORBmatcher.cc:42-221): the candidate
// lists of the left and the right search come from the enumeration kernels above (one DevFrame / DevQueries /
// SearchScratch per camera, WITHOUT the static occupancy filter); this replays the serial MapPoint order on them.
struct FisheyeResolveArgs {
  int m, n_left, n_right;
  float nnratio;
  const uint8_t *in_left, *in_right;  // [m] the two searches a point takes part in (view flags, far gate, level_r != -1)
  const uint8_t* has_obs;            // [m]
  const uint8_t* occupied;           // [n_left + n_right] initial slot state
  const int32_t *left_to_right, *right_to_left;
  SearchScratch left, right;
  int32_t* assign;                   // [n_left + n_right]
  int32_t* nmatches;
};
void launch_search_resolve_fisheye(const FisheyeResolveArgs& A, cudaStream_t st);

struct DevKeyFrame {
  int n, n_levels;
  const orbx_kp* kps;
  const uint8_t* desc;
  const float* u_right;        // or NULL
  const uint8_t* has_mappoint;
  int n_nodes;
  const uint32_t* node_ids;
  const int32_t* offsets;
  const uint32_t* indices;
  const float* scale_factors;
  const float* level_sigma2;
};
struct TriArgs {
  DevKeyFrame k1, k2;
  float F12[9];
  float ep_x, ep_y;
  int only_stereo, coarse, check_orientation;
  int32_t* matches12;  // [k1.n]
  int32_t* nmatches;   // [1]
  int32_t* node_match; // [k1.n_nodes] scratch: index of the same node id in k2 or -1
};
void launch_triangulation(const TriArgs& A, cudaStream_t st);
// fill == false: off[k1.n + 1] = exclusive offsets of the per-feature candidate counts, *total = their sum; fill == true:
// the candidates themselves (at most cap)
void launch_triangulation_candidates(const TriArgs& A, int32_t* off, int32_t* total, int32_t* cand_idx2,
                                     int32_t* cand_dist, int cap, bool fill, cudaStream_t st);
// SearchForInitialization: serial replay over the candidate lists of run_search (Q = the level-0 keypoints of F1)
struct InitArgs {
  const orbx_kp* kps1;      // F1.mvKeysUn
  float nnratio;
  int check_orientation;
  int n2;
  int32_t* matches12;       // [Q.m]
  int32_t* matches21;       // [n2] scratch
  int32_t* matched_dist;    // [n2] scratch
  int32_t* events;          // [2 * Q.m] scratch
  int32_t* nmatches;        // [1]
};
void launch_init_resolve(const DevFrame& F2, const DevQueries& Q, const SearchScratch& S, const InitArgs& A,
                         cudaStream_t st);
// SearchByBoW(KeyFrame*, Frame&, ...): kf.has_mappoint = the KeyFrame features that carry a good MapPoint
struct BowArgs {
  DevKeyFrame kf, fr;
  float nnratio;
  int check_orientation;
  int n_left_f;         // >= 0: fr is a two-camera Frame (rows < n_left_f = left camera): left / right bests, :274-365
  int kf_kf;            // 1: the KeyFrame-KeyFrame form (:766-884): matches indexed by kf, fr.has_mappoint read, < TH_LOW
  int32_t* matches_f;   // [fr.n] KeyFrame feature index or -1 (kf_kf: [kf.n] index into fr or -1)
  uint8_t* matched2;    // [fr.n] scratch (kf_kf): vbMatched2
  int32_t* nmatches;    // [1]
  int32_t* node_match;  // [kf.n_nodes] scratch: index of the same node id in fr or -1
};
void launch_search_by_bow(const BowArgs& A, cudaStream_t st);
// The matching loop of ORBmatcher::Fuse: Q.max_level = nPredictedLevel, Q.u_right = ur of the projection
void launch_fuse_match(const DevFrame& F, const DevQueries& Q, const float* inv_level_sigma2, int chi2_gate,
                       int32_t* best_idx, int32_t* best_dist, cudaStream_t st);
// Frame::AssignFeaturesToGrid for `frames` keypoint arrays (kp_stride apart; counts from n_ptr[f] or n_fixed)
void launch_build_grid(const orbx_kp* kps, const int32_t* n_ptr, int n_fixed, int64_t kp_stride, int frames, float min_x,
                       float min_y, float inv_w, float inv_h, int32_t* offsets, int32_t* items, int64_t item_stride,
                       cudaStream_t st);

}  // namespace orbx

#endif
