/* orbm.h — C ABI of the Hamming searches of the ORB front-end (liborbx.so).
 *
 * Replaces, behind unchanged C++ signatures (shim/), the data-parallel part of ORB_SLAM3::ORBmatcher
 * (include/ORBmatcher.h:38-133), Frame::ComputeStereoMatches (src/Frame.cc:921-1084) and the brute-force
 * cv::BFMatcher::knnMatch call of Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1293). The C++ shim flattens the
 * reference's pointer graph (MapPoint*, KeyFrame*, DBoW2::FeatureVector) into the SoA / CSR views of orbx_types.h and
 * scatters the results back.
 *
 * Conventions as in orbx.h. A matcher context owns a CUDA stream and grow-only device scratch; ORBmatcher objects are
 * per-call temporaries used concurrently from the Tracking / LocalMapping / LoopClosing threads
 * (src/Tracking.cc:2784, src/LocalMapping.cc:435), so keep one context per thread.
 */
#ifndef ORBM_H_
#define ORBM_H_

#include <stdint.h>

#include "orbx.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORBM_TH_LOW 50        /* ORBmatcher::TH_LOW       src/ORBmatcher.cc:36 */
#define ORBM_TH_HIGH 100      /* ORBmatcher::TH_HIGH      src/ORBmatcher.cc:35 */
#define ORBM_HISTO_LENGTH 30  /* ORBmatcher::HISTO_LENGTH src/ORBmatcher.cc:37 */

typedef struct orbm_matcher orbm_matcher;

int orbm_create(orbm_matcher** out, int device);
void orbm_destroy(orbm_matcher* m);
const char* orbm_last_error(const orbm_matcher* m);

/* static int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) (src/ORBmatcher.cc:1959-1973) for
 * n pairs at once: a[n][32], b[n][32] -> dist[n] (host buffers). The single-pair form is an inline popcount in the
 * shim header; a device launch per pair would be absurd. */
int orbm_descriptor_distance_batch(orbm_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist);

/* void MapPoint::ComputeDistinctiveDescriptors() (include/MapPoint.h:84, src/MapPoint.cc:372-441), the arithmetic after
 * the observed descriptors have been gathered (:407-435), batched over map points: descriptors
 * of point p are rows [offsets[p], offsets[p + 1]) of desc (host, 32 bytes each). best_idx[p] = position inside the
 * point's list of the descriptor with the least median Hamming distance to the rest (first one on ties), -1 for an
 * empty list; the shim then does mDescriptor = vDescriptors[best].clone(). */
int orbm_distinctive_descriptors(orbm_matcher* m, const uint8_t* desc, const int32_t* offsets, int n_points,
                                 int32_t* best_idx);

/* cv::BFMatcher(cv::NORM_HAMMING).knnMatch(query, train, matches, 2) (src/Frame.cc:1293): for every query row the two
 * train rows minimising (distance, trainIdx). idx = -1 and dist = -1 where the train set has fewer than 1 / 2 rows.
 * Host buffers: q[nq][32], t[nt][32], outputs [nq]. */
int orbm_knn2(orbm_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx1, int32_t* d1,
              int32_t* idx2, int32_t* d2);
/* Same with every pointer in device memory; enqueued on cuda_stream (NULL = the context's stream), not synchronised.
 * Stream contract of every *_device entry point of this header: the call uses the matcher context's scratch (partial
 * results, expanded descriptors, candidate slabs), so one context serves ONE stream at a time — device calls of the same
 * context on different streams must be ordered by the caller (see orbx_extract_batch_device in orbx.h). Contexts are
 * cheap; give every concurrent stream its own, as every thread has its own (threading). */
int orbm_knn2_device(orbm_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int32_t* d_idx1,
                     int32_t* d_d1, int32_t* d_idx2, int32_t* d_d2, void* cuda_stream);

/* void Frame::ComputeStereoMatches() (src/Frame.cc:921-1084). `left` / `right` are the extractors whose LAST call
 * produced the pyramids of this pair (mpORBextractorLeft/Right->mvImagePyramid, :927,1011,1029); `frame` selects the
 * frame of that call. kps / desc are mvKeys, mDescriptors, mvKeysRight, mDescriptorsRight (host). mbf, mb as in Frame.
 * Outputs (host): u_right[n_l] = mvuRight, depth[n_l] = mvDepth (-1 = no match); *n_matched = surviving matches.
 * Where the reference would index out of range (empty match list at :1073, windows outside the level) the pair is
 * simply left unmatched. */
int orbm_stereo_match(orbm_matcher* m, const orbx_extractor* left, const orbx_extractor* right, int frame,
                      const orbx_kp* kps_l, const uint8_t* desc_l, int n_l, const orbx_kp* kps_r,
                      const uint8_t* desc_r, int n_r, float mbf, float mb, float* u_right, float* depth,
                      int32_t* n_matched);
/* Batched, device-resident: pair p uses frame p of both extractors' last call and rows [p][cap] of the keypoint /
 * descriptor arrays with counts d_n_l[p], d_n_r[p] (exactly what orbx_extract_batch_device wrote). Outputs
 * d_u_right[p][cap], d_depth[p][cap], d_n_matched[p]. Enqueued on cuda_stream, not synchronised. */
int orbm_stereo_match_batch_device(orbm_matcher* m, const orbx_extractor* left, const orbx_extractor* right,
                                   int n_pairs, const orbx_kp* d_kps_l, const uint8_t* d_desc_l, const int32_t* d_n_l,
                                   const orbx_kp* d_kps_r, const uint8_t* d_desc_r, const int32_t* d_n_r, int cap,
                                   float mbf, float mb, float* d_u_right, float* d_depth, int32_t* d_n_matched,
                                   void* cuda_stream);

/* The hot path of the stereo Frame constructor (src/Frame.cc:149-279) for n_pairs independent rectified pairs, host
 * buffers in and out: both ORBextractor::operator() calls with vLappingArea = {0, 0} (:200-203) and
 * ComputeStereoMatches (:223). Pairs are cut into groups of max_batch and pipelined over the extractors' two lanes:
 * the H2D copy of group i+1 and the D2H copy of group i-1 overlap the kernels of group i (use orbx_host_alloc memory).
 * Outputs: kps / desc [n_pairs][cap] per eye with counts n_l / n_r, u_right / depth [n_pairs][cap], n_matched. */
int orbm_stereo_frames_batch(orbm_matcher* m, orbx_extractor* left, orbx_extractor* right, int n_pairs,
                             const uint8_t* imgs_l, const uint8_t* imgs_r, int width, int height, int stride,
                             int64_t frame_stride, float mbf, float mb, orbx_kp* kps_l, uint8_t* desc_l,
                             int32_t* n_l, orbx_kp* kps_r, uint8_t* desc_r, int32_t* n_r, int cap, float* u_right,
                             float* depth, int32_t* n_matched);

/* int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th,
 * const bool bFarPoints, const float thFarPoints) (include/ORBmatcher.h:49-51, src/ORBmatcher.cc:42-221), serial
 * MapPoint order, Nleft == -1. mfNNratio is the matcher's constructor argument (include/ORBmatcher.h:38).
 * assign[f->n] (host): index of the MapPoint written to F.mvpMapPoints[i], or -1. *nmatches = the return value. */
int orbm_search_by_projection_map(orbm_matcher* m, const orbx_frame_view* f, const orbx_mappoints* mps, float th,
                                  float nnratio, int far_points, float th_far, int32_t* assign, int32_t* nmatches);

/* The same function on a two-camera Frame (Nleft != -1; KannalaBrandt8 rigs, src/ORBmatcher.cc:42-221 with its
 * right-camera twin :148-217): per MapPoint the left search, then the right search on mGridRight / mvKeysRight, every
 * accepted point also written to the stereo partner of its keypoint (mvLeftToRightMatch / mvRightToLeftMatch). The
 * projections of both cameras come from the caller (Frame::isInFrustumChecks stays on the host: mpCamera2 is a
 * KannalaBrandt8 model). assign[f->n_left + f->n_right] (host): index of the MapPoint each row of F.mvpMapPoints ends up
 * with, or -1; *nmatches = the return value. Serial MapPoint order, exact. */
int orbm_search_by_projection_map_fisheye(orbm_matcher* m, const orbx_fisheye_view* f, const orbx_mappoints* mps,
                                          const orbx_mappoints_right* mr, float th, float nnratio, int far_points,
                                          float th_far, int32_t* assign, int32_t* nmatches);

/* void Frame::AssignFeaturesToGrid() with bool Frame::PosInGrid(const cv::KeyPoint&, int&, int&) (include/Frame.h:107,
 * 263; src/Frame.cc:520-547, 833-844), Nleft == -1: the 64x48 lookup grid of mvKeysUn as CSR (orbx_grid layout: cell
 * id = col * 48 + row, ascending keypoint indices inside a cell). SURVEY.md §8(f) rank 1 — the step immediately
 * before SearchByProjection. kps[n] (host) -> cell_offsets[64*48 + 1], cell_items[n] (host; the first
 * cell_offsets[64*48] entries are used). */
int orbm_assign_features_to_grid(orbm_matcher* m, const orbx_kp* kps, int n, float min_x, float min_y, float inv_w,
                                 float inv_h, int32_t* cell_offsets, int32_t* cell_items);

/* orbm_search_by_projection_map on a frame that never left the device: frame `frame` of the extractor's most recent
 * host-facing call (orbx_extract / orbx_extract_batch; n = the keypoint count that call returned). Keypoints and
 * descriptors are read from the extractor's device outputs, the lookup grid is built on the device
 * (Frame::AssignFeaturesToGrid, as above) and mvScaleFactors come from the extractor. Valid for an undistorted camera,
 * where Frame::UndistortKeyPoints leaves mvKeysUn = mvKeys (src/Frame.cc:562-571). u_right[n] (mvuRight) and
 * occupied[n] are host arrays and may be NULL (monocular / no keypoint holds a MapPoint yet). */
int orbm_search_by_projection_map_resident(orbm_matcher* m, const orbx_extractor* ex, int frame, int n,
                                           const float* u_right, const uint8_t* occupied, float min_x, float min_y,
                                           float inv_w, float inv_h, const orbx_mappoints* mps, float th, float nnratio,
                                           int far_points, float th_far, int32_t* assign, int32_t* nmatches);

/* bool Frame::isInFrustum(MapPoint* pMP, float viewingCosLimit) (include/Frame.h:101, src/Frame.cc:632-699, Nleft == -1)
 * with MapPoint::PredictScale (src/MapPoint.cc:559-573) for every point of local map `map_index`, as the loop of
 * Tracking::SearchLocalPoints runs it (src/Tracking.cc:3288-3300) — SURVEY.md §8(f) rank 1. Host buffers. Outputs
 * [map->m]: track_in_view = mbTrackInView (0 for skip[] points, which are not projected); proj_x / proj_y = mTrackProjX /
 * mTrackProjY (-1 when the point falls outside the image, :635-636); proj_xr, level, view_cos, depth = mTrackProjXR,
 * mnTrackScaleLevel, mTrackViewCos, mTrackDepth, meaningful only where track_in_view != 0 (the reference leaves them
 * untouched otherwise). *n_in_view = nToMatch. The arrays are exactly what orbx_mappoints takes. */
int orbm_is_in_frustum(orbm_matcher* m, const orbx_frustum* fr, const orbx_local_map* map, int map_index,
                       float viewing_cos_limit, uint8_t* track_in_view, float* proj_x, float* proj_y, float* proj_xr,
                       int32_t* level, float* view_cos, float* depth, int32_t* n_in_view);

/* void Tracking::SearchLocalPoints() (src/Tracking.cc:3249-3330) for n_frames frames of ONE extract batch, everything
 * resident in device memory and nothing synchronised (BASELINE.json configs[3] as a throughput path): per frame
 * isInFrustum over its local map, Frame::AssignFeaturesToGrid (src/Frame.cc:520-547) and
 * ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
 * (src/ORBmatcher.cc:42-221) in serial MapPoint order. Frame f = rows [f][cap] of d_kps / d_desc with count d_n[f]
 * (what orbx_extract_batch_device wrote; mvKeysUn = mvKeys: undistorted camera, src/Frame.cc:562-571), d_u_right /
 * d_occupied [f][cap] (mvuRight from orbm_stereo_match_batch_device; either may be NULL), pose d_frustums[f], local map
 * d_map_index[f] (NULL: f % maps->n_maps) of *maps, whose array pointers are DEVICE pointers. `ex` supplies
 * mvScaleFactors. Outputs: d_assign[f][cap] = index of the MapPoint written to mvpMapPoints[i] or -1, d_nmatches[f] =
 * SearchByProjection's return value, d_n_in_view[f] = nToMatch, d_status[f] = 0 or ORBX_E_CAPACITY (candidate list,
 * see orbx_track_params). cap < 65536. */
int orbm_track_local_map_batch_device(orbm_matcher* m, const orbx_extractor* ex, int n_frames, const orbx_kp* d_kps,
                                      const uint8_t* d_desc, const int32_t* d_n, int cap, const float* d_u_right,
                                      const uint8_t* d_occupied, const orbx_frustum* d_frustums,
                                      const orbx_local_map* maps, const int32_t* d_map_index,
                                      const orbx_track_params* prm, int32_t* d_assign, int32_t* d_nmatches,
                                      int32_t* d_n_in_view, int32_t* d_status, void* cuda_stream);

/* orbm_stereo_frames_batch followed, per pair, by Tracking::SearchLocalPoints on the left frame (configs[3] end to end):
 * host buffers in and out, pipelined over the lanes like orbm_stereo_frames_batch. Additional inputs (host):
 * frustums[n_pairs], the local maps (*maps with HOST pointers; uploaded once per call), map_index[n_pairs] or NULL
 * (pair p uses map p % n_maps), occupied[n_pairs][cap] or NULL. Additional outputs: assign[n_pairs][cap],
 * nmatches[n_pairs], n_in_view[n_pairs]. Returns ORBX_E_CAPACITY if any frame overflowed an output or the candidate
 * list. */
int orbm_stereo_track_frames_batch(orbm_matcher* m, orbx_extractor* left, orbx_extractor* right, int n_pairs,
                                   const uint8_t* imgs_l, const uint8_t* imgs_r, int width, int height, int stride,
                                   int64_t frame_stride, float mbf, float mb, const orbx_frustum* frustums,
                                   const orbx_local_map* maps, const int32_t* map_index, const uint8_t* occupied,
                                   const orbx_track_params* prm, orbx_kp* kps_l, uint8_t* desc_l, int32_t* n_l,
                                   orbx_kp* kps_r, uint8_t* desc_r, int32_t* n_r, int cap, float* u_right, float* depth,
                                   int32_t* n_matched, int32_t* assign, int32_t* nmatches, int32_t* n_in_view);

/* orbm_stereo_frames_batch / orbm_stereo_track_frames_batch over several GPUs of one box from ONE C/C++ caller: m[d],
 * left[d], right[d] live on the same device d' (any ordinals). Pairs are independent: they are cut into n_devices
 * contiguous blocks (as orbx_extract_batch_multi) and every block runs the single-device call on its own host thread;
 * the eyes of a pair never split, so ComputeStereoMatches and the local-map search stay local; no inter-GPU traffic.
 * The local maps (track form) are uploaded to every device. Pass frustums == NULL for the stereo-only form. */
int orbm_stereo_track_frames_batch_multi(int n_devices, orbm_matcher* const* m, orbx_extractor* const* left,
                                         orbx_extractor* const* right, int n_pairs, const uint8_t* imgs_l,
                                         const uint8_t* imgs_r, int width, int height, int stride, int64_t frame_stride,
                                         float mbf, float mb, const orbx_frustum* frustums, const orbx_local_map* maps,
                                         const int32_t* map_index, const uint8_t* occupied, const orbx_track_params* prm,
                                         orbx_kp* kps_l, uint8_t* desc_l, int32_t* n_l, orbx_kp* kps_r, uint8_t* desc_r,
                                         int32_t* n_r, int cap, float* u_right, float* depth, int32_t* n_matched,
                                         int32_t* assign, int32_t* nmatches, int32_t* n_in_view);

/* int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
 * (src/ORBmatcher.cc:1594-1806) and (Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist) (:1808-1918), after the
 * caller-side SE3 projection (orbx_projected). max_dist = TH_HIGH or ORBdist; check_orientation = mbCheckOrientation.
 * assign[f->n] (host): index into `pts` or -1. */
int orbm_search_by_projection_frame(orbm_matcher* m, const orbx_frame_view* f, const orbx_projected* pts,
                                    int max_dist, int check_orientation, int32_t* assign, int32_t* nmatches);

/* The candidate loop of the same function for ONE camera of a two-camera frame (CurrentFrame.Nleft != -1,
 * src/ORBmatcher.cc:1649-1690 on mvKeys / mGrid / rows [0, Nleft) and its twin :1711-1755 on mvKeysRight / mGridRight /
 * rows [Nleft, N)): f is that camera's view (u_right = NULL: the stereo gate of :1667 only exists for Nleft == -1).
 * Both cameras share ONE rotation histogram (:1693-1706, :1757-1778, :1785-1803), so no orientation check happens
 * here: decisions[pts->m] (host) = the camera-local keypoint each point is written to, or -1, under the serial order
 * dependence of :1663-1665 (a keypoint is closed by an earlier point with observations); the caller bins and filters.
 * window_count[pts->m] (host, may be NULL) = |GetFeaturesInArea(...)| of the point's window: the reference skips the
 * right-camera search of a point whose LEFT window is empty (:1655). shim/ORBmatcher_orbx.cc shows the composition. */
int orbm_search_by_projection_frame_decisions(orbm_matcher* m, const orbx_frame_view* f, const orbx_projected* pts,
                                              int max_dist, int32_t* decisions, int32_t* window_count);

/* orbm_search_by_projection_frame on a frame that never left the device (see orbm_search_by_projection_map_resident
 * for the conditions): SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, ...) right after ExtractORB,
 * src/Tracking.cc:2811. */
int orbm_search_by_projection_frame_resident(orbm_matcher* m, const orbx_extractor* ex, int frame, int n,
                                             const float* u_right, const uint8_t* occupied, float min_x, float min_y,
                                             float inv_w, float inv_h, const orbx_projected* pts, int max_dist,
                                             int check_orientation, int32_t* assign, int32_t* nmatches);

/* void Frame::ComputeBoW() / KeyFrame::ComputeBoW() (src/Frame.cc:846-851, src/KeyFrame.cc) — SURVEY.md §8(f) rank 2:
 * mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4). orbm_set_vocabulary uploads the DBoW2 tree once
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h: m_nodes, m_L; ORBvoc has 1.1 M nodes = 35 MB of descriptors) and
 * keeps it on the device; orbm_bow_transform runs the per-feature tree descent of transform(feature, word_id, weight,
 * &nid, levelsup) (:1218-1262) for n descriptors: word_id[n], weight[n] (the leaf's WordValue), node_id[n] (the node
 * at level m_L - levelsup, 0 = root). The shim then fills the two std::maps in feature order exactly as :1147-1160 does
 * (v.addWeight(id, w) when w > 0, fv.addFeature(nid, i)) and normalises (:1198), so every double is summed by the
 * reference's own code. */
int orbm_set_vocabulary(orbm_matcher* m, const orbx_vocabulary* voc);
int orbm_bow_transform(orbm_matcher* m, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight,
                       uint32_t* node_id);

/* int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched,
 * vector<int>& vnMatches12, int windowSize = 10) (include/ORBmatcher.h:73-77, src/ORBmatcher.cc:618-764) — SURVEY.md
 * §8(f) rank 3 — in the serial order of its loop (the fork's tbb::parallel_for at :634 races on vnMatches21 /
 * vMatchedDistance; the serial order is the well-defined semantics, as for the other searches). f1: F1.mvKeysUn and
 * F1.mDescriptors (its grid is not read); f2: F2 with grid; prev_matched_xy[f1->n][2] = vbPrevMatched.
 * matches12[f1->n] (host) = vnMatches12; the shim then refreshes vbPrevMatched from it (:758-761). */
int orbm_search_for_initialization(orbm_matcher* m, const orbx_frame_view* f1, const orbx_frame_view* f2,
                                   const float* prev_matched_xy, int window_size, float nnratio, int check_orientation,
                                   int32_t* matches12, int32_t* nmatches);

/* The matching loop of int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th,
 * const bool bRight) (include/ORBmatcher.h:94, src/ORBmatcher.cc:1108-1275; bRight = false, NLeft == -1) — SURVEY.md
 * §8(f) rank 3. The shim projects the points that pass :1141-1187 (orbx_projected: u, v, u_right = ur, radius =
 * th * mvScaleFactors[nPredictedLevel], max_level = nPredictedLevel, desc = GetDescriptor(); min_level, angle, has_obs
 * are not read); kf carries mvKeysUn, mDescriptors, mvuRight and the KeyFrame's grid; inv_level_sigma2 =
 * mvInvLevelSigma2[kf->n_levels]. best_idx[m] / best_dist[m] (host): the most similar keypoint that passes the level
 * window and the chi-square gate (:1194-1257), -1 / 256 when none. The shim then applies bestDist <= TH_LOW and does
 * the Replace / AddObservation surgery in point order (:1261-1273). chi2_gate = 1 for this overload; 0 gives the loop
 * of Fuse(KeyFrame*, Sophus::Sim3f& Scw, const vector<MapPoint*>&, float th, vector<MapPoint*>& vpReplacePoint)
 * (:1277-1390), which has the same window and level test but no reprojection gate (:1356-1372); inv_level_sigma2 is
 * not read then and may be NULL.
 * Two-camera KeyFrames (NLeft != -1) and bRight (:1116-1124, :1200-1201, :1219-1221, :1247): the search runs on ONE
 * camera, so kf is that camera's view — mvKeys / mGrid / descriptor rows [0, NLeft), or with bRight mvKeysRight /
 * mGridRight / rows [NLeft, N) with u_right = mvuRight under the camera-local index — and the caller adds NLeft to
 * best_idx (shim/ORBmatcher_next_orbx.cc; the projection stays with the reference's own mpCamera2 object). */
int orbm_fuse_match(orbm_matcher* m, const orbx_frame_view* kf, const float* inv_level_sigma2,
                    const orbx_projected* pts, int chi2_gate, int32_t* best_idx, int32_t* best_dist);

/* int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (include/ORBmatcher.h:66,
 * src/ORBmatcher.cc:230-404), Nleft == -1. kf->has_mappoint[i] = vpMapPointsKF[i] != NULL &&
 * !isBad(); `frame` carries F.mvKeys, F.mDescriptors and F.mFeatVec (its has_mappoint is not read; every feature may
 * appear under one node only, as DBoW2 produces — otherwise ORBX_E_ARG). mfNNratio / mbCheckOrientation are the
 * matcher's constructor arguments. matches_f[frame->n] (host) = index of the KeyFrame feature whose MapPoint goes to
 * vpMapPointMatches[i], or -1; *nmatches = the return value. */
int orbm_search_by_bow(orbm_matcher* m, const orbx_keyframe_view* kf, const orbx_keyframe_view* frame, float nnratio,
                       int check_orientation, int32_t* matches_f, int32_t* nmatches);

/* The same function on a two-camera Frame (F.Nleft != -1, src/ORBmatcher.cc:274-365): rows [0, n_left_frame) of `frame`
 * are the left camera's, the rest the right camera's; per KeyFrame feature the best two distances are kept apart for the
 * two cameras, the left best is accepted by the ratio test and — inside "bestDist1 <= TH_LOW" (:319) — the right best
 * whenever it is <= TH_LOW (its ratio test is switched off by "|| true", :347-350). kps of both views: left rows then
 * right rows (mvKeys / mvKeysRight, :323-335; mvKeysUn for a one-camera KeyFrame). Outputs as orbm_search_by_bow. */
int orbm_search_by_bow_fisheye(orbm_matcher* m, const orbx_keyframe_view* kf, const orbx_keyframe_view* frame,
                               int n_left_frame, float nnratio, int check_orientation, int32_t* matches_f,
                               int32_t* nmatches);

/* int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (include/ORBmatcher.h:67,
 * src/ORBmatcher.cc:766-884), NLeft == -1. has_mappoint[i] = vpMapPoints[i] != NULL && !isBad() on both sides;
 * matches12[kf1->n] (host) = index of the KeyFrame-2 feature whose MapPoint goes to vpMatches12[i], or -1. */
int orbm_search_by_bow_kf(orbm_matcher* m, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, float nnratio,
                          int check_orientation, int32_t* matches12, int32_t* nmatches);

/* int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vector<pair<size_t,size_t>>& vMatchedPairs,
 * const bool bOnlyStereo, const bool bCoarse) (src/ORBmatcher.cc:886-1106), pinhole keyframes. F12 (row-major 3x3)
 * and the epipole (ep_x, ep_y) are computed by the shim exactly as the reference does (:893-911,
 * src/CameraModels/Pinhole.cpp:122-149). matches12[kf1->n] (host) = index in kf2 or -1; the shim turns it into
 * vMatchedPairs in ascending idx1 order (:1097-1103). */
int orbm_search_for_triangulation(orbm_matcher* m, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2,
                                  const float* F12, float ep_x, float ep_y, int only_stereo, int coarse,
                                  int check_orientation, int32_t* matches12, int32_t* nmatches);

/* The descriptor part of the same function for two-camera rigs (both KeyFrames with mpCamera2; src/ORBmatcher.cc:
 * 973-988): the epipolar test of a KannalaBrandt8 pair is KannalaBrandt8::TriangulateMatches (camera models are the
 * caller's), so the device enumerates, per kf1 feature without a MapPoint, the kf2 features without a MapPoint under
 * the same vocabulary node whose descriptor distance is <= TH_LOW, in the reference's scan order, and the shim replays
 * ":988-1052" over them with the reference's own camera objects (shim/ORBmatcher_orbx.cc). offsets[kf1->n + 1] (host):
 * CSR offsets; cand_idx2 / cand_dist[cap] (host); *total = number of candidates. ORBX_E_CAPACITY when total > cap (the
 * first cap are written; call again with a larger buffer). u_right of the views is not read. */
int orbm_triangulation_candidates(orbm_matcher* m, const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2,
                                  int32_t* offsets, int32_t* cand_idx2, int32_t* cand_dist, int32_t cap, int32_t* total);

#ifdef __cplusplus
}
#endif
#endif /* ORBM_H_ */
