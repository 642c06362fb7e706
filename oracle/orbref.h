/* orbref.h — CPU ORACLE for the ORB front-end hot path. TEST INFRASTRUCTURE ONLY.
 *
 * A serial-order restatement of the reference's CPU algorithm (hellovuong/ORB_SLAM3_FAST @ 6255e16), used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker and as the
 * timed CPU baseline. Nothing under orb_slam3_fast_b200/ may include, link or load this library.
 *
 * PINNING STATUS
 *   - The reference has no tests, golden vectors or fixtures for this path, and its library cannot be
 *     built here as a whole (OpenCV C++ headers/libs, TBB, Eigen, Sophus, Pangolin: none installed, no network).
 *   - EXTRACTOR (src/ORBextractor.cc, rows a1-a10): pinned by the reference's OWN source. oracle/Makefile (target `ref`)
 *     compiles /root/reference/src/ORBextractor.cc where it lies into oracle/_ref/liborbref_src.so against stand-in
 *     OpenCV / TBB headers (oracle/ref_stubs: types with OpenCV's semantics, TBB run serially, the five image
 *     primitives forwarded to the cv2-pinned restatements below); tests/test_oracle_vs_reference_source.py requires
 *     this oracle to equal it bit for bit (constructor tables, 9 image / size / lapping cases, parameter variants,
 *     degenerate images, and DistributeOctTree called directly on 65 k-candidate levels).
 *   - The third-party arithmetic the reference calls (OpenCV resize / GaussianBlur / FAST / copyMakeBorder /
 *     fastAtan2 / BFMatcher, glibc cosf/sinf, libstdc++ std::sort) is pinned: tests/test_oracle_primitives.py checks
 *     every primitive below byte-for-byte against the real OpenCV 4.13.0 kernels through cv2, and
 *     tests/test_oracle_pipeline.py checks the whole extractor against an independent Python pipeline that drives the
 *     real cv2 kernels (oracle/cv2_pipeline.py), from which tests/golden/ is frozen.
 *   - MATCHERS of src/ORBmatcher.cc (rows a11, a13-a15 and the §8f searches): pinned by the reference's OWN source too.
 *     oracle/_ref/liborbref_matcher_src.so is /root/reference/src/ORBmatcher.cc compiled where it lies against the
 *     stand-in Frame / KeyFrame / MapPoint world of ref_stubs/matcher_world.h (their real headers need Eigen, Sophus,
 *     g2o, boost); tests/test_oracle_matchers_vs_reference_source.py requires the restatements below to equal it word
 *     for word (DescriptorDistance, SearchByProjection x3, SearchForTriangulation, SearchByBoW x2,
 *     SearchForInitialization, Fuse x2).
 *   - DBoW2's transform (orbref_bow_transform): pinned by DBoW2's own TemplatedVocabulary / FORB code as vendored by the
 *     reference, compiled in place into oracle/_ref/liborbref_dbow2_src.so (tests/test_oracle_vs_reference_source.py).
 *   - Frame::ComputeStereoMatches (src/Frame.cc:921-1084, row a12): pinned by the reference's own text. Frame.cc cannot
 *     be compiled as a whole, so oracle/Makefile cuts this one function out of the reference file by its signature and
 *     pipes it to the compiler inside the stand-in world (nothing but the object file is written); the test runs the
 *     reference's operator() on both images + its ComputeStereoMatches and requires mvuRight / mvDepth to equal
 *     orbref_stereo_match bit for bit.
 *   - Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea, KeyFrame::GetFeaturesInArea and
 *     MapPoint::ComputeDistinctiveDescriptors: piped in the same way and checked directly.
 *   - Frame::isInFrustum + MapPoint::PredictScale (orbref_is_in_frustum): the reference's own text piped in the same way,
 *     against a stand-in Eigen whose 3-term reductions follow Eigen's own order c0 + (c1 + c2) (no Eigen headers exist in
 *     this image, so THAT ORDER is this repository's reading of Eigen/src/Core/Redux.h, not something verified here) and
 *     a stand-in Pinhole::project with the reference's expression. knn2 is pinned to
 *     cv2.BFMatcher itself. No function below is left without a reference-source (or, for the OpenCV primitives, cv2) pin.
 *
 * Build: oracle/Makefile (g++ -O2, no -march=native, -ffp-contract=off: the reference is built without FMA,
 * CMakeLists.txt:13-18).
 */
#ifndef ORBREF_H_
#define ORBREF_H_

#include "../include/orbx_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- OpenCV / glibc primitives (restated; pinned to cv2 4.13.0 in tests) ---- */
void orbref_resize_linear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orbref_gauss7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
void orbref_border101(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int border);
/* cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16. Returns the count; fills up to cap entries. */
int orbref_fast9(const uint8_t* img, int w, int h, int stride, int threshold, int* xs, int* ys, int* scores, int cap);
float orbref_fast_atan2(float y, float x);
/* How the image primitives above were built (reported by bench.py next to the CPU baseline). */
const char* orbref_primitives_kind(void);
int orbref_cv_round(float v);
/* libstdc++ std::sort on (key0, key1) pairs with the reference's compareNodes ordering (src/ORBextractor.cc:542-555);
 * writes the resulting permutation (perm[i] = original index of the element now at position i). */
void orbref_std_sort_perm(const int* key0, const int* key1, int n, int* perm);

/* ---- ORBextractor (src/ORBextractor.cc, serial path) ---- */
typedef struct orbref_extractor orbref_extractor;
orbref_extractor* orbref_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void orbref_extractor_destroy(orbref_extractor* ex);
/* tables: each array has nlevels entries (umax has 16). Any pointer may be NULL. */
void orbref_extractor_tables(const orbref_extractor* ex, float* scale, float* inv_scale, float* sigma2,
                             float* inv_sigma2, int* features_per_level, int* umax);
/* operator(): returns 0, or -1 on empty image, or -2 if cap is too small. *mono_index = the reference's return. */
int orbref_extract(orbref_extractor* ex, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                   orbx_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_index);
/* intermediates of the last orbref_extract call (for stage-by-stage parity tests) */
int orbref_level_dims(const orbref_extractor* ex, int level, int* w, int* h);
const uint8_t* orbref_level_image(const orbref_extractor* ex, int level, int* stride);    /* ROI of the bordered buffer */
const uint8_t* orbref_level_bordered(const orbref_extractor* ex, int level, int* stride); /* (w+38) x (h+38) */
const uint8_t* orbref_level_blurred(const orbref_extractor* ex, int level, int* stride);
int orbref_level_candidates(const orbref_extractor* ex, int level, orbx_kp* out, int cap); /* before the quadtree */
int orbref_level_keypoints(const orbref_extractor* ex, int level, orbx_kp* out, int cap);  /* after, with angle */

/* ---- matching ---- */
int orbref_descriptor_distance(const uint8_t* a, const uint8_t* b); /* src/ORBmatcher.cc:1959-1973 */
/* cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) (src/Frame.cc:1293): per query the two train rows minimising
 * (distance, trainIdx). idx = -1 / dist = -1 when the train set has fewer rows. */
void orbref_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx1, int32_t* d1, int32_t* idx2,
                 int32_t* d2);
/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:372-441), the arithmetic after the observations have been
 * gathered: all-pairs distances of the n descriptors, per row the element [0.5 * (n - 1)] of the sorted row, first row
 * with the least such median. Returns the index of that descriptor, -1 for n == 0. */
int orbref_distinctive_descriptor(const uint8_t* desc, int n);
/* Frame::ComputeStereoMatches (src/Frame.cc:921-1084). Reads the raw pyramids of the two extractors' last extract.
 * best_dist_out (optional) receives the SAD of accepted matches (or -1). Returns the number of surviving matches. */
int orbref_stereo_match(const orbref_extractor* left, const orbref_extractor* right, const orbx_kp* kps_l,
                        const uint8_t* desc_l, int n_l, const orbx_kp* kps_r, const uint8_t* desc_r, int n_r,
                        float mbf, float mb, float* u_right, float* depth);
/* Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:520-547,833-844). offsets[64*48+1], items[n]. */
void orbref_build_grid(const orbx_kp* kps, int n, float min_x, float min_y, float inv_w, float inv_h,
                       int32_t* offsets, int32_t* items);
/* bool Frame::isInFrustum(MapPoint*, viewingCosLimit) (src/Frame.cc:632-699, Nleft == -1) over the points of local map
 * `map_index`, as the loop of Tracking::SearchLocalPoints runs it (src/Tracking.cc:3288-3300): a point with skip[] set is
 * not projected (track_in_view = 0, nothing else written); for the others track_in_view = the return value, proj_x /
 * proj_y = -1 or the projection (:635-636, :655-656), and the remaining MapPoint tracking fields (:676-685) are written
 * only when the point is in view. Outputs have map->m entries. */
void orbref_is_in_frustum(const orbx_frustum* fr, const orbx_local_map* map, int map_index, float viewing_cos_limit,
                          uint8_t* track_in_view, float* proj_x, float* proj_y, float* proj_xr, int32_t* level,
                          float* view_cos, float* depth);
/* Frame::GetFeaturesInArea (src/Frame.cc:765-831), Nleft == -1. Returns the count written to out (cap >= n). */
int orbref_features_in_area(const orbx_frame_view* f, float x, float y, float r, int min_level, int max_level,
                            int32_t* out);
/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints), serial MapPoint
 * order (src/ORBmatcher.cc:42-221), Nleft == -1. assign[n]: index of the MapPoint written to each keypoint or -1
 * (keypoints that were already occupied stay -1). Returns nmatches. */
int orbref_search_by_projection_map(const orbx_frame_view* f, const orbx_mappoints* mps, float th, float nnratio,
                                    int far_points, float th_far, int32_t* assign);
/* The same function on a two-camera Frame (Nleft != -1): per MapPoint, in vector order, the left search (:61-145 with
 * the Nleft != -1 choices: no uRight gate, octave from mvKeys, the accepted point is also written to the stereo partner
 * mvLeftToRightMatch[bestIdx] + Nleft, +1 match) and then the right-camera twin (:148-217: window on mTrackProjXR / YR at
 * mnTrackScaleLevelR with NO th factor, mGridRight, rows idx + Nleft, the partner mvRightToLeftMatch[bestIdx]). A slot is
 * closed for a later search while its CURRENT occupant has observations (:92-93, :183-185); partner writes are
 * unconditional. mps->track_in_view / mr->track_in_view_r already include !isBad(); mps->proj_xr is not read.
 * assign[n_left + n_right]: index of the MapPoint each slot ends up with, or -1. Returns nmatches. */
int orbref_search_by_projection_map_fisheye(const orbx_fisheye_view* f, const orbx_mappoints* mps,
                                            const orbx_mappoints_right* mr, float th, float nnratio, int far_points,
                                            float th_far, int32_t* assign);
/* ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) / (Frame&, KeyFrame*, set, th, ORBdist) after the
 * caller-side projection (src/ORBmatcher.cc:1594-1806, 1808-1918). block_any != 0 selects the keyframe variant's
 * "any MapPoint blocks" rule (:1862). assign[n] as above. Returns nmatches. */
int orbref_search_by_projection_frame(const orbx_frame_view* f, const orbx_projected* pts, int max_dist,
                                      int check_orientation, int32_t* assign);
int orbref_search_by_projection_frame_decisions(const orbx_frame_view* f, const orbx_projected* pts, int max_dist,
                                                int32_t* decisions, int32_t* window);
/* cv::remap(src, dst, mapx, mapy, INTER_LINEAR) with CV_32FC1 maps, 8-bit single channel, BORDER_CONSTANT 0 — the stereo
 * rectification of System::TrackStereo (src/System.cc:293-294; maps from initUndistortRectifyMap(..., CV_32F, ...),
 * src/Settings.cc:557-572). OpenCV's fixed point: coordinates rounded to 1/32 px (cvRound(map * 32)), the four taps
 * weighted with 5-bit fractions, (sum + 512) >> 10 (pinned to cv2 4.13 in tests/test_oracle_primitives.py). */
void orbref_remap_linear(const uint8_t* src, int sw, int sh, int sstride, const float* mapx, const float* mapy, int dw,
                         int dh, uint8_t* dst, int dstride);
/* cv::cvtColor(src, dst, COLOR_{BGR,RGB,BGRA,RGBA}2GRAY) on 8-bit images, the conversion Tracking::GrabImage* applies to
 * colour input (src/Tracking.cc:1394-1412, 1500-1513, 1558-1571): OpenCV's fixed point,
 * gray = (B * 3735 + G * 19235 + R * 9798 + 16384) >> 15 (pinned to cv2 4.13 in tests/test_oracle_primitives.py).
 * channels = 3 or 4; rgb != 0 when the first channel is red. */
void orbref_cvt_gray(const uint8_t* src, int w, int h, int stride, int channels, int rgb, uint8_t* dst, int dstride);
/* The per-feature part of Frame::ComputeBoW (src/Frame.cc:846-851) = TemplatedVocabulary::transform(feature, word_id,
 * weight, &nid, levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1262): descend the tree taking at every
 * level the FIRST child with the least FORB::distance; nid = the node passed at level m_L - levelsup (0 = root when that
 * level is <= 0 or never reached). Outputs per feature. */
void orbref_bow_transform(const orbx_vocabulary* voc, const uint8_t* desc, int n, int levelsup, uint32_t* word_id,
                          double* weight, uint32_t* node_id);
/* ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12,
 * int windowSize) (src/ORBmatcher.cc:618-764), in the SERIAL order of its loop (the fork wraps it in a racy
 * tbb::parallel_for, :634). f1: mvKeysUn + mDescriptors of F1 (grid unused); f2: F2 with its grid; prev_xy[f1->n][2] =
 * vbPrevMatched. matches12[f1->n] = vnMatches12. Returns nmatches. (The caller then refreshes vbPrevMatched, :758-761.) */
int orbref_search_for_initialization(const orbx_frame_view* f1, const orbx_frame_view* f2, const float* prev_xy,
                                     int window_size, float nnratio, int check_orientation, int32_t* matches12);
/* The matching part of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight = false)
 * (src/ORBmatcher.cc:1108-1275), NLeft == -1, after the caller-side projection (:1152-1192): per point
 * KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:705-749), the level window [L-1, L] (:1221), the chi-square gate on the
 * reprojection error (7.8 stereo / 5.99 mono, :1223-1244), the most similar keypoint (:1250-1257).
 * kf: mvKeysUn, mDescriptors, mvuRight, grid of the KeyFrame. pts: u, v, u_right (ur), radius, max_level =
 * nPredictedLevel (min_level, angle, has_obs unused), desc = MapPoint::GetDescriptor(). best_idx[i] = keypoint or -1,
 * best_dist[i] = its distance (256 when none); the caller applies bestDist <= TH_LOW and the graph surgery.
 * chi2_gate = 0 gives the loop of Fuse(KeyFrame*, Sophus::Sim3f&, ...) (:1356-1372), which has no such gate. */
void orbref_fuse_match(const orbx_frame_view* kf, const float* inv_level_sigma2, const orbx_projected* pts,
                       int chi2_gate, int32_t* best_idx, int32_t* best_dist);
/* ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (src/ORBmatcher.cc:766-884),
 * NLeft == -1. has_mappoint[i] = vpMapPoints[i] != NULL && !isBad() on both sides. matches12[kf1->n] = index of the
 * KeyFrame-2 feature whose MapPoint is written to vpMatches12[i], or -1. Returns nmatches. */
int orbref_search_by_bow_kf(const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, float nnratio,
                            int check_orientation, int32_t* matches12);
/* ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:230-404),
 * Nleft == -1. kf->has_mappoint[i] = vpMapPointsKF[i] != NULL && !isBad(); `frame` carries F.mvKeys, F.mDescriptors and
 * F.mFeatVec (has_mappoint unused). matches_f[frame->n] = index of the KeyFrame feature whose MapPoint is written to
 * vpMapPointMatches[i], or -1. Returns nmatches. */
int orbref_search_by_bow(const orbx_keyframe_view* kf, const orbx_keyframe_view* frame, float nnratio,
                         int check_orientation, int32_t* matches_f);
/* ... on a two-camera Frame (Nleft != -1): left / right bests kept apart, :274-365. */
int orbref_search_by_bow_fisheye(const orbx_keyframe_view* kf, const orbx_keyframe_view* frame, int n_left_f,
                                 float nnratio, int check_orientation, int32_t* matches_f);
/* ORBmatcher::SearchForTriangulation, pinhole mono/stereo keyframes (src/ORBmatcher.cc:886-1106 +
 * src/CameraModels/Pinhole.cpp:122-149). F12 = K1^-T [t12]x R12 K2^-1 computed by the caller (row-major 3x3);
 * ep = projection of camera centre 1 into image 2. matches12[kf1->n] = idx2 or -1. Returns nmatches. */
int orbref_search_for_triangulation(const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, const float* F12,
                                    float ep_x, float ep_y, int only_stereo, int coarse, int check_orientation,
                                    int32_t* matches12);
/* ... on two-camera KeyFrames (src/ORBmatcher.cc:886-1106 with its NLeft / mpCamera2 branches): F12x4 = four row-major
 * 3x3 matrices indexed 2 * right1 + right2, the Pinhole-form epipolar test of the selected camera pair. */
int orbref_search_for_triangulation_fisheye(const orbx_keyframe_view* kf1, int n_left1, const orbx_keyframe_view* kf2,
                                            int n_left2, const float* F12x4, int only_stereo, int coarse,
                                            int check_orientation, int32_t* matches12);
/* The descriptor part of SearchForTriangulation (:973-988): per kf1 feature the kf2 candidates with distance <= TH_LOW
 * in scan order (CSR); returns the total, writes at most cap entries. */
int orbref_triangulation_candidates(const orbx_keyframe_view* kf1, const orbx_keyframe_view* kf2, int32_t* offsets,
                                    int32_t* cand_idx2, int32_t* cand_dist, int cap);

#ifdef __cplusplus
}
#endif
#endif
