// C entry points over the reference's own ORB_SLAM3::ORBmatcher (compiled from /root/reference/src/ORBmatcher.cc against
// the stand-in world of oracle/ref_stubs/matcher_world.h). Same flat views and result conventions as the oracle's
// functions of the same name (oracle/orbref.h), so a test compares the two outputs directly. TEST INFRASTRUCTURE.
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "ORBmatcher.h"

using namespace ORB_SLAM3;

namespace {
cv::Mat rows32(const uint8_t* p, int n) { return cv::Mat(n, 32, CV_8UC1, const_cast<uint8_t*>(p), 32); }

std::vector<cv::KeyPoint> keypoints(const orbx_kp* k, int n) {
  std::vector<cv::KeyPoint> v(n);
  if (n) memcpy(static_cast<void*>(v.data()), k, (size_t)n * sizeof(cv::KeyPoint));
  return v;
}

void fill_common(FeatureHolder& h, const orbx_kp* kps, const uint8_t* desc, const float* u_right, int n,
                 const float* scale_factors, const float* level_sigma2, int n_levels) {
  h.N = n;
  h.Nleft = h.NLeft = -1;
  h.mvKeys = h.mvKeysUn = keypoints(kps, n);
  h.mvuRight.assign(n, -1.f);
  if (u_right) h.mvuRight.assign(u_right, u_right + n);
  h.mDescriptors = rows32(desc, n);
  if (scale_factors) h.mvScaleFactors.assign(scale_factors, scale_factors + n_levels);
  if (level_sigma2) h.mvLevelSigma2.assign(level_sigma2, level_sigma2 + n_levels);
}

void fill_featvec(DBoW2::FeatureVector& fv, const orbx_featvec& c) {
  for (int a = 0; a < c.n_nodes; a++)
    fv[c.node_ids[a]] = std::vector<unsigned int>(c.indices + c.offsets[a], c.indices + c.offsets[a + 1]);
}

inline void set_static(float& s, float v) {
  if (s != v) s = v;
}

// Frame::mGrid built by the reference's AssignFeaturesToGrid from mvKeysUn and the view's bounds / cell sizes
void build_grid(Frame& f, const orbx_frame_view* v) {
  // the image bounds / cell sizes are statics (as in the reference, include/Frame.h:372-378): written only when they
  // change, so that threads working on frames of one geometry (the three-thread tests) only ever read them
  set_static(Frame::mnMinX, v->grid.min_x);
  set_static(Frame::mnMinY, v->grid.min_y);
  set_static(Frame::mfGridElementWidthInv, v->grid.inv_w);
  set_static(Frame::mfGridElementHeightInv, v->grid.inv_h);
  f.AssignFeaturesToGrid();
}

// a KeyFrame over a frame view: features from the view, grid copied from a Frame as KeyFrame::KeyFrame does
void keyframe_from_view(KeyFrame& kf, const orbx_frame_view* v) {
  fill_common(kf, v->kps, v->desc, v->u_right, v->n, v->scale_factors, nullptr, v->n_levels);
  Frame tmp;
  fill_common(tmp, v->kps, v->desc, v->u_right, v->n, v->scale_factors, nullptr, v->n_levels);
  build_grid(tmp, v);
  kf.take_grid(tmp);
}

// a KeyFrame whose feature i holds MapPoint &points[i] when has_mappoint[i]
struct KeyFrameWorld {
  KeyFrame kf;
  std::vector<MapPoint> points;
  GeometricCamera camera;
  explicit KeyFrameWorld(const orbx_keyframe_view* v) : points(v->n) {
    fill_common(kf, v->kps, v->desc, v->u_right, v->n, v->scale_factors, v->level_sigma2, v->n_levels);
    fill_featvec(kf.mFeatVec, v->featvec);
    kf.mvpMapPoints.assign(v->n, nullptr);
    for (int i = 0; i < v->n; i++)
      if (v->has_mappoint && v->has_mappoint[i]) kf.mvpMapPoints[i] = &points[i];
    kf.mpCamera = &camera;
  }
  int index_of(const MapPoint* p) const { return p ? (int)(p - points.data()) : -1; }
};

struct FrameWorld {
  Frame f;
  MapPoint occupied;  // stands for "a MapPoint with observations" on the keypoints flagged occupied
  GeometricCamera camera;
  explicit FrameWorld(const orbx_frame_view* v) {
    fill_common(f, v->kps, v->desc, v->u_right, v->n, v->scale_factors, nullptr, v->n_levels);
    build_grid(f, v);
    occupied.observations = 1;
    f.mvpMapPoints.assign(v->n, nullptr);
    for (int i = 0; i < v->n; i++)
      if (v->occupied && v->occupied[i]) f.mvpMapPoints[i] = &occupied;
    f.mpCamera = &camera;
  }
};
}  // namespace

// cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2) for the stand-in OpenCV: the oracle's knn2 (pinned to cv2.BFMatcher by
// tests/test_oracle_primitives.py) reshaped into OpenCV's output — up to k matches per query, best first.
void cv::BFMatcher::knnMatch(const cv::Mat& query, const cv::Mat& train, std::vector<std::vector<cv::DMatch>>& matches,
                             int k) const {
  const int nq = query.rows, nt = train.rows;
  std::vector<uint8_t> q((size_t)std::max(nq, 1) * 32), t((size_t)std::max(nt, 1) * 32);
  for (int i = 0; i < nq; i++) memcpy(&q[(size_t)i * 32], query.ptr(i), 32);
  for (int i = 0; i < nt; i++) memcpy(&t[(size_t)i * 32], train.ptr(i), 32);
  std::vector<int32_t> i1(std::max(nq, 1)), d1(std::max(nq, 1)), i2(std::max(nq, 1)), d2(std::max(nq, 1));
  orbref_knn2(q.data(), nq, t.data(), nt, i1.data(), d1.data(), i2.data(), d2.data());
  matches.assign(nq, std::vector<cv::DMatch>());
  for (int i = 0; i < nq; i++) {
    if (k >= 1 && i1[i] >= 0) {
      cv::DMatch m;
      m.queryIdx = i; m.trainIdx = i1[i]; m.distance = (float)d1[i];
      matches[i].push_back(m);
    }
    if (k >= 2 && i2[i] >= 0) {
      cv::DMatch m;
      m.queryIdx = i; m.trainIdx = i2[i]; m.distance = (float)d2[i];
      matches[i].push_back(m);
    }
  }
}

extern "C" {

int orbrefsrc_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  return ORBmatcher::DescriptorDistance(rows32(a, 1), rows32(b, 1));
}

int orbrefsrc_search_by_projection_map(const orbx_frame_view* fv, const orbx_mappoints* mps, float th, float nnratio,
                                       int far_points, float th_far, int32_t* assign) {
  FrameWorld W(fv);
  std::vector<MapPoint> pts(mps->m);
  std::vector<MapPoint*> ptrs(mps->m);
  for (int i = 0; i < mps->m; i++) {
    MapPoint& p = pts[i];
    p.mbTrackInView = mps->track_in_view[i] != 0;
    p.mTrackProjX = mps->proj_x[i];
    p.mTrackProjY = mps->proj_y[i];
    p.mTrackProjXR = mps->proj_xr ? mps->proj_xr[i] : 0.f;
    p.mnTrackScaleLevel = mps->level[i];
    p.mTrackViewCos = mps->view_cos[i];
    p.mTrackDepth = mps->depth[i];
    p.observations = mps->has_obs[i] ? 1 : 0;
    p.descriptor = rows32(mps->desc + (size_t)i * 32, 1);
    ptrs[i] = &p;
  }
  ORBmatcher matcher(nnratio, true);
  const int n = matcher.SearchByProjection(W.f, ptrs, th, far_points != 0, th_far);
  for (int i = 0; i < fv->n; i++) {
    const MapPoint* p = W.f.mvpMapPoints[i];
    assign[i] = (p && p != &W.occupied) ? (int)(p - pts.data()) : -1;
  }
  return n;
}

int orbrefsrc_search_for_triangulation(const orbx_keyframe_view* v1, const orbx_keyframe_view* v2, const float* F12,
                                       float ep_x, float ep_y, int only_stereo, int coarse, int check_orientation,
                                       int32_t* matches12) {
  KeyFrameWorld A(v1), B(v2);
  memcpy(A.camera.F12, F12, sizeof(A.camera.F12));
  // camera centre 1 at (ep_x, ep_y, 0), keyframe 2 at the origin, orthographic stand-in camera: ep = (ep_x, ep_y)
  A.kf.pose = Sophus::SE3f(Eigen::Matrix3f(), Eigen::Vector3f(-ep_x, -ep_y, 0.f));
  ORBmatcher matcher(0.6f, check_orientation != 0);
  std::vector<std::pair<size_t, size_t>> pairs;
  const int n = matcher.SearchForTriangulation(&A.kf, &B.kf, pairs, only_stereo != 0, coarse != 0);
  for (int i = 0; i < v1->n; i++) matches12[i] = -1;
  for (const auto& pr : pairs) matches12[pr.first] = (int32_t)pr.second;
  return n;
}

// ... on two-camera KeyFrames: NLeft set, rows >= NLeft in mvKeysRight, both cameras of both KeyFrames set. Camera 1 /
// camera 2 of KeyFrame 1 carry (F12, F12b) = the matrices for a (left | right) second feature; see matcher_world.h.
int orbrefsrc_search_for_triangulation_fisheye(const orbx_keyframe_view* v1, int n_left1, const orbx_keyframe_view* v2,
                                               int n_left2, const float* F12x4, int only_stereo, int coarse,
                                               int check_orientation, int32_t* matches12) {
  KeyFrameWorld A(v1), B(v2);
  GeometricCamera a2, b2;
  auto split = [&](KeyFrame& kf, const orbx_keyframe_view* v, int nl, GeometricCamera* second) {
    kf.Nleft = kf.NLeft = nl;
    kf.mvKeys = keypoints(v->kps, nl);
    kf.mvKeysRight = keypoints(v->kps + nl, v->n - nl);
    kf.mvKeysUn.clear();
    kf.mpCamera2 = second;
    second->tag = 1;
  };
  split(A.kf, v1, n_left1, &a2);
  split(B.kf, v2, n_left2, &b2);
  memcpy(A.camera.F12, F12x4, 36);        // left1  x left2
  memcpy(A.camera.F12b, F12x4 + 9, 36);   // left1  x right2
  memcpy(a2.F12, F12x4 + 18, 36);         // right1 x left2
  memcpy(a2.F12b, F12x4 + 27, 36);        // right1 x right2
  ORBmatcher matcher(0.6f, check_orientation != 0);
  std::vector<std::pair<size_t, size_t>> pairs;
  const int n = matcher.SearchForTriangulation(&A.kf, &B.kf, pairs, only_stereo != 0, coarse != 0);
  for (int i = 0; i < v1->n; i++) matches12[i] = -1;
  for (const auto& pr : pairs) matches12[pr.first] = (int32_t)pr.second;
  return n;
}

int orbrefsrc_search_by_bow(const orbx_keyframe_view* kfv, const orbx_keyframe_view* frame, float nnratio,
                            int check_orientation, int32_t* matches_f) {
  KeyFrameWorld K(kfv);
  Frame F;
  fill_common(F, frame->kps, frame->desc, frame->u_right, frame->n, frame->scale_factors, nullptr, frame->n_levels);
  fill_featvec(F.mFeatVec, frame->featvec);
  ORBmatcher matcher(nnratio, check_orientation != 0);
  std::vector<MapPoint*> out;
  const int n = matcher.SearchByBoW(&K.kf, F, out);
  for (int i = 0; i < frame->n; i++) matches_f[i] = K.index_of(out[i]);
  return n;
}

// ... on a two-camera rig: rows >= NLeft of a view are the right camera's keypoints (mvKeysRight), both mpCamera2 set —
// the fisheye branches of :274-365 run
int orbrefsrc_search_by_bow_fisheye(const orbx_keyframe_view* kfv, int n_left_kf, const orbx_keyframe_view* frame,
                                    int n_left_f, float nnratio, int check_orientation, int32_t* matches_f) {
  KeyFrameWorld K(kfv);
  GeometricCamera cam2;
  auto split = [&](FeatureHolder& h, const orbx_keyframe_view* v, int nl) {
    h.Nleft = h.NLeft = nl;
    h.mvKeys = keypoints(v->kps, nl);
    h.mvKeysRight = keypoints(v->kps + nl, v->n - nl);
    h.mvKeysUn.clear();  // must not be read in this mode
    h.mpCamera2 = &cam2;
  };
  split(K.kf, kfv, n_left_kf);
  Frame F;
  fill_common(F, frame->kps, frame->desc, frame->u_right, frame->n, frame->scale_factors, nullptr, frame->n_levels);
  fill_featvec(F.mFeatVec, frame->featvec);
  F.mpCamera = &cam2;
  split(F, frame, n_left_f);
  ORBmatcher matcher(nnratio, check_orientation != 0);
  std::vector<MapPoint*> out;
  const int n = matcher.SearchByBoW(&K.kf, F, out);
  for (int i = 0; i < frame->n; i++) matches_f[i] = K.index_of(out[i]);
  return n;
}

int orbrefsrc_search_by_bow_kf(const orbx_keyframe_view* v1, const orbx_keyframe_view* v2, float nnratio,
                               int check_orientation, int32_t* matches12) {
  KeyFrameWorld A(v1), B(v2);
  ORBmatcher matcher(nnratio, check_orientation != 0);
  std::vector<MapPoint*> out;
  const int n = matcher.SearchByBoW(&A.kf, &B.kf, out);
  for (int i = 0; i < v1->n; i++) matches12[i] = B.index_of(out[i]);
  return n;
}

// ... on two-camera KeyFrames (:799-801, :816-818): NLeft != -1 and mvKeysUn holds the first n_un rows only — rows past
// it (the right camera's) are skipped on both sides
int orbrefsrc_search_by_bow_kf_fisheye(const orbx_keyframe_view* v1, int n_un1, const orbx_keyframe_view* v2, int n_un2,
                                       float nnratio, int check_orientation, int32_t* matches12) {
  KeyFrameWorld A(v1), B(v2);
  A.kf.NLeft = A.kf.Nleft = n_un1;
  A.kf.mvKeysUn.resize(n_un1);
  B.kf.NLeft = B.kf.Nleft = n_un2;
  B.kf.mvKeysUn.resize(n_un2);
  ORBmatcher matcher(nnratio, check_orientation != 0);
  std::vector<MapPoint*> out;
  const int n = matcher.SearchByBoW(&A.kf, &B.kf, out);
  for (int i = 0; i < v1->n; i++) matches12[i] = B.index_of(out[i]);
  return n;
}

int orbrefsrc_search_for_initialization(const orbx_frame_view* f1, const orbx_frame_view* f2, const float* prev_xy,
                                        int window_size, float nnratio, int check_orientation, int32_t* matches12) {
  FrameWorld A(f1), B(f2);
  std::vector<cv::Point2f> prev(f1->n);
  for (int i = 0; i < f1->n; i++) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
  ORBmatcher matcher(nnratio, check_orientation != 0);
  std::vector<int> m12;
  const int n = matcher.SearchForInitialization(A.f, B.f, prev, m12, window_size);
  for (int i = 0; i < f1->n; i++) matches12[i] = m12[i];
  return n;
}

// SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono = false), :1594-1806. The points of the
// last frame are placed at (u, v, z) in front of an identity-pose current frame with the orthographic stand-in camera,
// so x3Dc = (u, v, z) and uv = (u, v) exactly; mode = +1 / -1 / 0 puts the last frame in front of / behind / beside the
// current one (bForward / bBackward / neither). octave, angle = LastFrame.mvKeys[i]; has_obs = Observations() > 0.
int orbrefsrc_search_by_projection_last_frame(const orbx_frame_view* fv, int m, const float* u, const float* v,
                                              const float* z, const int32_t* octave, const float* angle,
                                              const uint8_t* has_obs, const uint8_t* desc, float th, float mbf, float mb,
                                              int mode, int check_orientation, int32_t* assign) {
  FrameWorld W(fv);
  W.f.mbf = mbf;
  W.f.mb = mb;
  W.f.mnMaxX = W.f.mnMaxY = 1e9f;  // (mnMinX / mnMinY stay the grid origin)
  Frame last;
  last.N = m;
  last.Nleft = -1;
  last.mvKeys.resize(m);
  last.mvKeysUn.resize(m);
  last.mvbOutlier.assign(m, false);
  std::vector<MapPoint> pts(m);
  last.mvpMapPoints.resize(m);
  for (int i = 0; i < m; i++) {
    last.mvKeys[i].octave = last.mvKeysUn[i].octave = octave[i];
    last.mvKeys[i].angle = last.mvKeysUn[i].angle = angle[i];
    pts[i].pos = Eigen::Vector3f(u[i], v[i], z[i]);
    pts[i].observations = has_obs[i] ? 1 : 0;
    pts[i].descriptor = rows32(desc + (size_t)i * 32, 1);
    last.mvpMapPoints[i] = &pts[i];
  }
  last.pose = Sophus::SE3f(Eigen::Matrix3f(), Eigen::Vector3f(0.f, 0.f, mode > 0 ? 1.f : (mode < 0 ? -1.f : 0.f)));
  ORBmatcher matcher(0.9f, check_orientation != 0);
  const int n = matcher.SearchByProjection(W.f, last, th, false);
  for (int i = 0; i < fv->n; i++) {
    const MapPoint* p = W.f.mvpMapPoints[i];
    assign[i] = (p && p != &W.occupied) ? (int)(p - pts.data()) : -1;
  }
  return n;
}

// ... on a two-camera current frame (Nleft != -1): the left search (:1649-1706) and, for the points whose left window
// held a feature, the right-camera twin (:1708-1780); GetRelativePoseTrl is the identity in the stand-in world, so the
// right search runs at the same (u, v) on mvKeysRight / mGridRight. last_right[i] != 0 makes last-frame feature i a
// right-camera one (rows >= LastFrame.Nleft). assign[n_left + n_right] = the point written to each row, or -1.
int orbrefsrc_search_by_projection_last_frame_fisheye(const orbx_fisheye_view* fv, int m, const float* u, const float* v,
                                                      const float* z, const int32_t* octave, const float* angle,
                                                      const uint8_t* has_obs, const uint8_t* desc, float th, float mbf,
                                                      float mb, int mode, int check_orientation, int32_t* assign) {
  Frame F;
  const int NL = fv->n_left, NR = fv->n_right, N = NL + NR;
  F.N = N;
  F.Nleft = NL;
  F.mvKeys = keypoints(fv->kps_left, NL);
  F.mvKeysRight = keypoints(fv->kps_right, NR);
  F.mvKeysUn = F.mvKeys;
  F.mDescriptors = rows32(fv->desc, N);
  F.mvuRight.assign(N, -1.f);
  F.mvScaleFactors.assign(fv->scale_factors, fv->scale_factors + fv->n_levels);
  Frame::mnMinX = fv->grid_left.min_x;
  Frame::mnMinY = fv->grid_left.min_y;
  Frame::mfGridElementWidthInv = fv->grid_left.inv_w;
  Frame::mfGridElementHeightInv = fv->grid_left.inv_h;
  F.AssignFeaturesToGrid();
  F.mbf = mbf;
  F.mb = mb;
  F.mnMaxX = F.mnMaxY = 1e9f;
  MapPoint occupied;
  occupied.observations = 1;
  F.mvpMapPoints.assign(N, nullptr);
  for (int i = 0; i < N; i++)
    if (fv->occupied && fv->occupied[i]) F.mvpMapPoints[i] = &occupied;
  GeometricCamera camera;
  F.mpCamera = F.mpCamera2 = &camera;
  // the last frame: the first half of its features are left-camera ones, the rest right-camera ones
  Frame last;
  last.N = m;
  last.Nleft = m / 2;
  last.mvKeys.resize(last.Nleft);
  last.mvKeysRight.resize(m - last.Nleft);
  last.mvKeysUn.resize(m);
  last.mvbOutlier.assign(m, false);
  std::vector<MapPoint> pts(m);
  last.mvpMapPoints.resize(m);
  for (int i = 0; i < m; i++) {
    cv::KeyPoint& kp = i < last.Nleft ? last.mvKeys[i] : last.mvKeysRight[i - last.Nleft];
    kp.octave = octave[i];
    kp.angle = angle[i];
    last.mvKeysUn[i].octave = -7;  // must not be read on a two-camera last frame
    last.mvKeysUn[i].angle = -7.f;
    pts[i].pos = Eigen::Vector3f(u[i], v[i], z[i]);
    pts[i].observations = has_obs[i] ? 1 : 0;
    pts[i].descriptor = rows32(desc + (size_t)i * 32, 1);
    last.mvpMapPoints[i] = &pts[i];
  }
  last.pose = Sophus::SE3f(Eigen::Matrix3f(), Eigen::Vector3f(0.f, 0.f, mode > 0 ? 1.f : (mode < 0 ? -1.f : 0.f)));
  ORBmatcher matcher(0.9f, check_orientation != 0);
  const int n = matcher.SearchByProjection(F, last, th, false);
  for (int i = 0; i < N; i++) {
    const MapPoint* p = F.mvpMapPoints[i];
    assign[i] = (p && p != &occupied) ? (int)(p - pts.data()) : -1;
  }
  return n;
}

// SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist), :1808-1918.
// Same placement; level = what MapPoint::PredictScale returns; found[i] puts point i into sAlreadyFound. Every keypoint
// flagged occupied in fv holds a MapPoint (here any MapPoint blocks, :1862).
int orbrefsrc_search_by_projection_keyframe(const orbx_frame_view* fv, int m, const float* u, const float* v,
                                            const int32_t* level, const float* angle, const uint8_t* found,
                                            const uint8_t* desc, float th, int orb_dist, int check_orientation,
                                            int32_t* assign) {
  FrameWorld W(fv);
  W.f.mnMaxX = W.f.mnMaxY = 1e9f;  // (mnMinX / mnMinY stay the grid origin)
  KeyFrame kf;
  kf.N = m;
  kf.mvKeysUn.resize(m);
  std::vector<MapPoint> pts(m);
  kf.mvpMapPoints.resize(m);
  std::set<MapPoint*> already;
  for (int i = 0; i < m; i++) {
    kf.mvKeysUn[i].angle = angle[i];
    pts[i].pos = Eigen::Vector3f(u[i], v[i], 1.f);
    pts[i].predicted_level = level[i];
    pts[i].observations = 1;
    pts[i].descriptor = rows32(desc + (size_t)i * 32, 1);
    kf.mvpMapPoints[i] = &pts[i];
    if (found[i]) already.insert(&pts[i]);
  }
  ORBmatcher matcher(0.9f, check_orientation != 0);
  const int n = matcher.SearchByProjection(W.f, &kf, already, th, orb_dist);
  for (int i = 0; i < fv->n; i++) {
    const MapPoint* p = W.f.mvpMapPoints[i];
    assign[i] = (p && p != &W.occupied) ? (int)(p - pts.data()) : -1;
  }
  return n;
}

// Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight = false), :1108-1281 (sim3 == 0) and
// Fuse(KeyFrame*, Sophus::Sim3f&, const vector<MapPoint*>&, th, vpReplacePoint), :1283-1390 (sim3 != 0). The KeyFrame sits
// at the origin with the orthographic stand-in camera and holds no MapPoints, point i is placed at (u, v, z) with its
// normal along the viewing ray, PredictScale returns level[i]. best_idx[i] = the keypoint the point was fused into
// (read off AddObservation) or -1 when the reference left it alone (bestDist > TH_LOW or no candidate).
int orbrefsrc_fuse(const orbx_frame_view* kfv, const float* inv_level_sigma2, int m, const float* u, const float* v,
                   const float* z, const int32_t* level, const uint8_t* desc, float th, float mbf, int sim3,
                   int32_t* best_idx) {
  KeyFrame kf;
  keyframe_from_view(kf, kfv);
  kf.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + kfv->n_levels);
  kf.mbf = mbf;
  kf.mnMaxX = kf.mnMaxY = 1e9f;
  kf.mvpMapPoints.assign(kfv->n, nullptr);
  GeometricCamera camera;
  kf.mpCamera = &camera;
  std::vector<MapPoint> pts(m);
  std::vector<MapPoint*> ptrs(m);
  for (int i = 0; i < m; i++) {
    pts[i].pos = Eigen::Vector3f(u[i], v[i], z[i]);
    pts[i].normal = pts[i].pos * 10.f;
    pts[i].predicted_level = level[i];
    pts[i].descriptor = rows32(desc + (size_t)i * 32, 1);
    ptrs[i] = &pts[i];
  }
  ORBmatcher matcher(0.6f, true);
  int n;
  if (sim3) {
    Sophus::Sim3f Scw;
    std::vector<MapPoint*> replace(m, nullptr);
    n = matcher.Fuse(&kf, Scw, ptrs, th, replace);
  } else {
    n = matcher.Fuse(&kf, ptrs, th, false);
  }
  for (int i = 0; i < m; i++) best_idx[i] = pts[i].added_to;
  return n;
}

// Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight) on a two-camera KeyFrame (NLeft != -1), :1108-1281 with the
// camera / pose / grid / keypoint selection of :1116-1124, :1200-1201, :1219-1221 and the "+ NLeft" of :1247: the left
// search (b_right == 0: mvKeys, mGrid, rows [0, NLeft)) or the right one (b_right != 0: mvKeysRight, mGridRight, rows
// [NLeft, N), pKF->mpCamera2, GetRightPose). Placement as orbrefsrc_fuse; both poses are the identity in the stand-in
// world. mvuRight is all -1 on such KeyFrames (Frame::ComputeStereoFishEyeMatches, src/Frame.cc:1281-1286), so only the
// 5.99 gate runs. best_idx[i] = the ROW of mDescriptors / mvpMapPoints point i was fused into, or -1.
int orbrefsrc_fuse_two_camera(const orbx_fisheye_view* fv, const float* inv_level_sigma2, int m, const float* u,
                              const float* v, const float* z, const int32_t* level, const uint8_t* desc, float th,
                              float mbf, int b_right, int32_t* best_idx) {
  const int NL = fv->n_left, NR = fv->n_right, N = NL + NR;
  Frame F;
  F.N = N;
  F.Nleft = NL;
  F.mvKeys = keypoints(fv->kps_left, NL);
  F.mvKeysRight = keypoints(fv->kps_right, NR);
  Frame::mnMinX = fv->grid_left.min_x;
  Frame::mnMinY = fv->grid_left.min_y;
  Frame::mfGridElementWidthInv = fv->grid_left.inv_w;
  Frame::mfGridElementHeightInv = fv->grid_left.inv_h;
  F.AssignFeaturesToGrid();
  KeyFrame kf;
  kf.N = N;
  kf.Nleft = kf.NLeft = NL;
  kf.mvKeys = F.mvKeys;
  kf.mvKeysRight = F.mvKeysRight;
  kf.mDescriptors = rows32(fv->desc, N);
  kf.mvuRight.assign(N, -1.f);
  kf.mvScaleFactors.assign(fv->scale_factors, fv->scale_factors + fv->n_levels);
  kf.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + fv->n_levels);
  kf.take_grid(F);
  kf.mbf = mbf;
  kf.mnMaxX = kf.mnMaxY = 1e9f;
  kf.mvpMapPoints.assign(N, nullptr);
  GeometricCamera camera, camera2;
  kf.mpCamera = &camera;
  kf.mpCamera2 = &camera2;
  std::vector<MapPoint> pts(m);
  std::vector<MapPoint*> ptrs(m);
  for (int i = 0; i < m; i++) {
    pts[i].pos = Eigen::Vector3f(u[i], v[i], z[i]);
    pts[i].normal = pts[i].pos * 10.f;
    pts[i].predicted_level = level[i];
    pts[i].descriptor = rows32(desc + (size_t)i * 32, 1);
    ptrs[i] = &pts[i];
  }
  ORBmatcher matcher(0.6f, true);
  const int n = matcher.Fuse(&kf, ptrs, th, b_right != 0);
  for (int i = 0; i < m; i++) best_idx[i] = pts[i].added_to;
  return n;
}

// SearchByProjection(KeyFrame*, Sophus::Sim3f& Scw, const vector<MapPoint*>&, vector<MapPoint*>& vpMatched, th,
// ratioHamming), :406-506 (with_kfs == 0) and the overload that also records the source KeyFrames, :508-616
// (with_kfs != 0). Identity Scw, same placement as orbrefsrc_fuse. matched_in[i] != 0: keypoint i already holds a match.
// assign[i] = index of the point written to vpMatched[i] by this call, or -1.
int orbrefsrc_search_by_projection_sim3(const orbx_frame_view* kfv, const uint8_t* matched_in, int m, const float* u,
                                        const float* v, const int32_t* level, const uint8_t* desc, int th,
                                        float ratio_hamming, int with_kfs, int32_t* assign) {
  KeyFrame kf;
  keyframe_from_view(kf, kfv);
  kf.mnMaxX = kf.mnMaxY = 1e9f;
  kf.mvpMapPoints.assign(kfv->n, nullptr);
  GeometricCamera camera;
  kf.mpCamera = &camera;
  MapPoint earlier;
  std::vector<MapPoint> pts(m);
  std::vector<MapPoint*> ptrs(m), matched(kfv->n, nullptr);
  for (int i = 0; i < kfv->n; i++)
    if (matched_in[i]) matched[i] = &earlier;
  for (int i = 0; i < m; i++) {
    pts[i].pos = Eigen::Vector3f(u[i], v[i], 1.f);
    pts[i].normal = pts[i].pos * 10.f;
    pts[i].predicted_level = level[i];
    pts[i].descriptor = rows32(desc + (size_t)i * 32, 1);
    ptrs[i] = &pts[i];
  }
  ORBmatcher matcher(0.75f, true);
  Sophus::Sim3f Scw;
  int n;
  if (with_kfs) {
    KeyFrame source;
    std::vector<KeyFrame*> kfs(m, &source), matched_kf(kfv->n, nullptr);
    n = matcher.SearchByProjection(&kf, Scw, ptrs, kfs, matched, matched_kf, th, ratio_hamming);
  } else {
    n = matcher.SearchByProjection(&kf, Scw, ptrs, matched, th, ratio_hamming);
  }
  for (int i = 0; i < kfv->n; i++) assign[i] = (matched[i] && matched[i] != &earlier) ? (int)(matched[i] - pts.data()) : -1;
  return n;
}

// SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const Sophus::Sim3f& S12, th),
// :1392-1592. Identity poses and S12, fx = fy = 1, cx = cy = 0: the MapPoint of feature i of KeyFrame k sits at
// (u_k[i], v_k[i], 1) and projects to exactly that in the other KeyFrame. has_k[i]: feature i holds a MapPoint;
// level_k[i] = what PredictScale returns for it; desc_k = MapPoint::GetDescriptor(). matches12[i1] = idx2 or -1.
int orbrefsrc_search_by_sim3(const orbx_frame_view* v1, const orbx_frame_view* v2, const uint8_t* has1, const float* u1,
                             const float* vv1, const int32_t* level1, const uint8_t* desc1, const uint8_t* has2,
                             const float* u2, const float* vv2, const int32_t* level2, const uint8_t* desc2, float th,
                             int32_t* matches12) {
  KeyFrame kf[2];
  std::vector<MapPoint> pts[2];
  const orbx_frame_view* views[2] = {v1, v2};
  const uint8_t* has[2] = {has1, has2};
  const float* us[2] = {u1, u2};
  const float* vs[2] = {vv1, vv2};
  const int32_t* levels[2] = {level1, level2};
  const uint8_t* descs[2] = {desc1, desc2};
  for (int k = 0; k < 2; k++) {
    const orbx_frame_view* v = views[k];
    keyframe_from_view(kf[k], v);
    kf[k].mnMaxX = kf[k].mnMaxY = 1e9f;
    pts[k].resize(v->n);
    kf[k].mvpMapPoints.assign(v->n, nullptr);
    for (int i = 0; i < v->n; i++) {
      if (!has[k][i]) continue;
      pts[k][i].pos = Eigen::Vector3f(us[k][i], vs[k][i], 1.f);
      pts[k][i].predicted_level = levels[k][i];
      pts[k][i].descriptor = rows32(descs[k] + (size_t)i * 32, 1);
      kf[k].mvpMapPoints[i] = &pts[k][i];
    }
  }
  ORBmatcher matcher(0.75f, true);
  std::vector<MapPoint*> m12(v1->n, nullptr);
  Sophus::Sim3f S12;
  const int n = matcher.SearchBySim3(&kf[0], &kf[1], m12, S12, th);
  for (int i = 0; i < v1->n; i++) matches12[i] = m12[i] ? (int)(m12[i] - pts[1].data()) : -1;
  return n;
}

// The stereo Frame constructor's hot path with the reference's own code end to end: ORBextractor::operator() on both
// images (vLappingArea = {0, 0}, src/Frame.cc:200-203), then Frame::ComputeStereoMatches (:921-1084, its text piped into
// this library at build time). Outputs: keypoints / descriptors of both eyes (cap rows), mvuRight / mvDepth [n_l].
int orbrefsrc_stereo_frame(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                           const unsigned char* img_l, const unsigned char* img_r, int w, int h, int stride, float mbf,
                           float mb, void* kps_l, unsigned char* desc_l, int* n_l, void* kps_r, unsigned char* desc_r,
                           int* n_r, float* u_right, float* depth, int cap) {
  ORBextractor left(nfeatures, scale_factor, nlevels, ini_th, min_th), right(nfeatures, scale_factor, nlevels, ini_th, min_th);
  Frame F;
  F.mpORBextractorLeft = &left;
  F.mpORBextractorRight = &right;
  std::vector<int> lapping = {0, 0};
  cv::Mat mask;
  cv::Mat il(h, w, CV_8UC1, const_cast<unsigned char*>(img_l), (size_t)stride);
  cv::Mat ir(h, w, CV_8UC1, const_cast<unsigned char*>(img_r), (size_t)stride);
  left(il, mask, F.mvKeys, F.mDescriptors, lapping);
  right(ir, mask, F.mvKeysRight, F.mDescriptorsRight, lapping);
  F.N = (int)F.mvKeys.size();
  F.mvScaleFactors = left.GetScaleFactors();
  F.mvInvScaleFactors = left.GetInverseScaleFactors();
  F.mbf = mbf;
  F.mb = mb;
  *n_l = F.N;
  *n_r = (int)F.mvKeysRight.size();
  if (*n_l > cap || *n_r > cap) return -1000;
  F.ComputeStereoMatches();
  if (*n_l) memcpy(kps_l, F.mvKeys.data(), (size_t)*n_l * sizeof(cv::KeyPoint));
  if (*n_r) memcpy(kps_r, F.mvKeysRight.data(), (size_t)*n_r * sizeof(cv::KeyPoint));
  for (int i = 0; i < *n_l; i++) memcpy(desc_l + (size_t)i * 32, F.mDescriptors.ptr(i), 32);
  for (int i = 0; i < *n_r; i++) memcpy(desc_r + (size_t)i * 32, F.mDescriptorsRight.ptr(i), 32);
  int matched = 0;
  for (int i = 0; i < *n_l; i++) {
    u_right[i] = F.mvuRight[i];
    depth[i] = F.mvDepth[i];
    matched += F.mvuRight[i] >= 0;
  }
  return matched;
}

// Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1271-1331): the brute-force 2-NN of the lapping-area descriptors,
// Lowe's ratio and the triangulation call, with the stand-in KannalaBrandt8 of matcher_world.h. In
// liborbref_matcher_src.so the function is the reference's own text; in libshim_world*.so it is the drop-in body of
// shim/FrameStereo_orbx.cc. Outputs: mvLeftToRightMatch[n_l], mvRightToLeftMatch[n_r], mvDepth[n_l],
// mvStereo3Dpoints[n_l] (xyz; untouched rows read 0); returns the number of accepted pairs.
int orbrefsrc_stereo_fisheye(const orbx_kp* kps_l, const unsigned char* desc_l, int n_l, const orbx_kp* kps_r,
                             const unsigned char* desc_r, int n_r, int mono_left, int mono_right,
                             const float* level_sigma2, int n_levels, const float* Rlr, const float* tlr,
                             int32_t* left_to_right, int32_t* right_to_left, float* depth, float* p3d) {
  Frame F;
  KannalaBrandt8 cam1, cam2;
  F.mpCamera = &cam1;
  F.mpCamera2 = &cam2;
  F.N = n_l + n_r;
  F.Nleft = n_l;
  F.Nright = n_r;
  F.monoLeft = mono_left;
  F.monoRight = mono_right;
  F.mvKeys = keypoints(kps_l, n_l);
  F.mvKeysRight = keypoints(kps_r, n_r);
  F.mDescriptors = rows32(desc_l, n_l);
  F.mDescriptorsRight = rows32(desc_r, n_r);
  F.mvLevelSigma2.assign(level_sigma2, level_sigma2 + n_levels);
  for (int i = 0; i < 9; i++) F.mRlr(i / 3, i % 3) = Rlr[i];
  F.mtlr = Eigen::Vector3f(tlr[0], tlr[1], tlr[2]);
  F.ComputeStereoFishEyeMatches();
  int matched = 0;
  for (int i = 0; i < n_l; i++) {
    left_to_right[i] = F.mvLeftToRightMatch[i];
    depth[i] = F.mvDepth[i];
    for (int c = 0; c < 3; c++) p3d[3 * i + c] = F.mvStereo3Dpoints[i](c);
    matched += F.mvLeftToRightMatch[i] >= 0;
  }
  for (int i = 0; i < n_r; i++) right_to_left[i] = F.mvRightToLeftMatch[i];
  return matched;
}

// Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:520-547, 833-844) as CSR in the oracle's cell order
// (cell = col * 48 + row): offsets[64 * 48 + 1], items[n].
void orbrefsrc_build_grid(const orbx_frame_view* v, int32_t* offsets, int32_t* items) {
  Frame f;
  fill_common(f, v->kps, v->desc, v->u_right, v->n, v->scale_factors, nullptr, v->n_levels);
  build_grid(f, v);
  int at = 0;
  for (int c = 0; c < FRAME_GRID_COLS; c++)
    for (int r = 0; r < FRAME_GRID_ROWS; r++) {
      offsets[c * FRAME_GRID_ROWS + r] = at;
      for (size_t i : f.mGrid[c][r]) items[at++] = (int32_t)i;
    }
  offsets[FRAME_GRID_COLS * FRAME_GRID_ROWS] = at;
}

// Frame::GetFeaturesInArea (:765-831) and, with keyframe != 0, KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:705-749,
// no level filter). Returns the count written to out.
int orbrefsrc_features_in_area(const orbx_frame_view* v, float x, float y, float r, int min_level, int max_level,
                               int keyframe, int32_t* out) {
  std::vector<size_t> idx;
  if (keyframe) {
    KeyFrame kf;
    keyframe_from_view(kf, v);
    idx = kf.GetFeaturesInArea(x, y, r);
  } else {
    FrameWorld W(v);
    idx = W.f.GetFeaturesInArea(x, y, r, min_level, max_level);
  }
  for (size_t i = 0; i < idx.size(); i++) out[i] = (int32_t)idx[i];
  return (int)idx.size();
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:372-441): one observing KeyFrame per descriptor row (the
// KeyFrames sit in one array, so the std::map<KeyFrame*, ...> visits them in row order). out[32] = mDescriptor.
// Returns 0 when the reference left mDescriptor empty (no observation).
int orbrefsrc_distinctive_descriptor(const uint8_t* desc, int n, uint8_t* out) {
  std::vector<KeyFrame> kfs(n);
  MapPoint mp;
  for (int i = 0; i < n; i++) {
    kfs[i].mDescriptors = rows32(desc + (size_t)i * 32, 1);
    mp.mObservations[&kfs[i]] = std::make_tuple(0, -1);
  }
  mp.ComputeDistinctiveDescriptors();
  if (mp.mDescriptor.empty()) return 0;
  memcpy(out, mp.mDescriptor.ptr(0), 32);
  return 1;
}
}

// Frame::isInFrustum (src/Frame.cc:632-699) + MapPoint::PredictScale (src/MapPoint.cc:559-573): the reference's own text
// on a stand-in Frame / MapPoint, driven like the loop of Tracking::SearchLocalPoints (src/Tracking.cc:3288-3300).
// Outputs as orbref_is_in_frustum.
extern "C" void orbrefsrc_is_in_frustum(const orbx_frustum* fr, const orbx_local_map* map, int map_index,
                                        float viewing_cos_limit, uint8_t* track_in_view, float* proj_x, float* proj_y,
                                        float* proj_xr, int32_t* level, float* view_cos, float* depth) {
  Frame F;
  PinholeStandIn cam;
  cam.mvParameters[0] = fr->fx; cam.mvParameters[1] = fr->fy; cam.mvParameters[2] = fr->cx; cam.mvParameters[3] = fr->cy;
  F.mpCamera = &cam;
  F.Nleft = -1;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) F.mRcw(i, j) = fr->Rcw[3 * i + j];
    F.mtcw(i) = fr->tcw[i];
    F.mOw(i) = fr->Ow[i];
  }
  F.mbf = fr->mbf;
  Frame::mnMinX = fr->min_x; Frame::mnMaxX = fr->max_x; Frame::mnMinY = fr->min_y; Frame::mnMaxY = fr->max_y;
  F.mfLogScaleFactor = fr->log_scale_factor;
  F.mnScaleLevels = fr->n_levels;
  const size_t base = (size_t)map_index * (size_t)map->m;
  for (int i = 0; i < map->m; i++) {
    const size_t g = base + (size_t)i;
    track_in_view[i] = 0;
    if (map->skip && map->skip[g]) continue;
    MapPoint p;
    p.pos = Eigen::Vector3f(map->pos[3 * g], map->pos[3 * g + 1], map->pos[3 * g + 2]);
    p.normal = Eigen::Vector3f(map->normal[3 * g], map->normal[3 * g + 1], map->normal[3 * g + 2]);
    p.use_raw_distances = true;
    p.mfMinDistance = map->min_dist[g];
    p.mfMaxDistance = map->max_dist[g];
    // every tracking field starts from the caller's value so that "left untouched" is observable
    p.mTrackProjXR = proj_xr[i]; p.mTrackDepth = depth[i]; p.mnTrackScaleLevel = level[i]; p.mTrackViewCos = view_cos[i];
    const bool in = F.isInFrustum(&p, viewing_cos_limit);
    track_in_view[i] = (in && p.mbTrackInView) ? 1 : 0;
    proj_x[i] = p.mTrackProjX;
    proj_y[i] = p.mTrackProjY;
    proj_xr[i] = p.mTrackProjXR;
    depth[i] = p.mTrackDepth;
    level[i] = p.mnTrackScaleLevel;
    view_cos[i] = p.mTrackViewCos;
  }
}

// void Tracking::SearchLocalPoints() (src/Tracking.cc:3249-3330) on a stand-in Tracking object: in
// liborbref_matcher_src.so the function is the reference's own text (cut out by its signature, oracle/Makefile), in the
// shim worlds it is the drop-in body of shim/Tracking_orbx.cc. The frame comes from a frame view (grid by the
// reference's AssignFeaturesToGrid), the pose from `fr`, mvpLocalMapPoints from local map `map_index`.
//   held[fv->n]   what mCurrentFrame.mvpMapPoints[i] holds on entry: -1 nothing, k >= 0 local-map point k, -2 a point
//                 with observations that is not in the local map
//   bad[m]        MapPoint::isBad()
//   ctl[8]        mSensor, isImuInitialized, GetIniertialBA2, mState, mCurrentFrame.mnId, mnLastRelocFrameId,
//                 mbFarPoints, two-camera flag (Nleft = N: isInFrustum takes its Nleft != -1 branch, whose
//                 isInFrustumChecks is a stand-in that sees nothing, so no point comes into view)
// In / out per local-map point (the caller's values are the state on entry, so "left untouched" is observable):
// track_in_view, proj_x, proj_y, proj_xr, level, view_cos, depth, visible (mnVisible), last_seen (mnLastFrameSeen).
// Out: assign[fv->n] = what mvpMapPoints[i] holds on return (same code as held), project_points[m][2] = the entry of
// mCurrentFrame.mmProjectPoints under the point's mnId (NaN, NaN when there is none). Returns mmProjectPoints.size().
extern "C" int orbrefsrc_search_local_points(const orbx_frame_view* fv, const orbx_frustum* fr, const orbx_local_map* map,
                                             int map_index, const int32_t* held, const uint8_t* bad, const int32_t* ctl,
                                             float th_far, int32_t* assign, uint8_t* track_in_view, float* proj_x,
                                             float* proj_y, float* proj_xr, int32_t* level, float* view_cos, float* depth,
                                             int32_t* visible, int32_t* last_seen, float* project_points) {
  Tracking T;
  Atlas atlas;
  LocalMapping mapper;
  PinholeStandIn cam;
  T.mpAtlas = &atlas;
  T.mpLocalMapper = &mapper;
  T.mSensor = ctl[0];
  atlas.imu_initialized = ctl[1] != 0;
  atlas.map.inertial_ba2 = ctl[2] != 0;
  T.mState = (Tracking::eTrackingState)ctl[3];
  T.mnLastRelocFrameId = (unsigned int)ctl[5];
  mapper.mbFarPoints = ctl[6] != 0;
  mapper.mThFarPoints = th_far;
  Frame& F = T.mCurrentFrame;
  fill_common(F, fv->kps, fv->desc, fv->u_right, fv->n, fv->scale_factors, nullptr, fv->n_levels);
  build_grid(F, fv);
  F.mnId = (long unsigned int)ctl[4];
  if (ctl[7]) F.Nleft = F.N;
  cam.mvParameters[0] = fr->fx; cam.mvParameters[1] = fr->fy; cam.mvParameters[2] = fr->cx; cam.mvParameters[3] = fr->cy;
  F.mpCamera = &cam;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) F.mRcw(i, j) = fr->Rcw[3 * i + j];
    F.mtcw(i) = fr->tcw[i];
    F.mOw(i) = fr->Ow[i];
  }
  F.pose = Sophus::SE3f(F.mRcw, F.mtcw);  // Frame::GetPose(); mRcw / mtcw are its parts (Frame::UpdatePoseMatrices)
  F.mbf = fr->mbf;
  set_static(Frame::mnMinX, fr->min_x); set_static(Frame::mnMaxX, fr->max_x); set_static(Frame::mnMinY, fr->min_y);
  set_static(Frame::mnMaxY, fr->max_y);
  F.mfLogScaleFactor = fr->log_scale_factor;
  F.mnScaleLevels = fr->n_levels;
  const int M = map->m;
  const size_t base = (size_t)map_index * (size_t)M;
  std::vector<MapPoint> pts(M);
  T.mvpLocalMapPoints.resize(M);
  for (int i = 0; i < M; i++) {
    const size_t g = base + (size_t)i;
    MapPoint& p = pts[i];
    p.mnId = 1000 + (long unsigned int)i;
    p.pos = Eigen::Vector3f(map->pos[3 * g], map->pos[3 * g + 1], map->pos[3 * g + 2]);
    p.normal = Eigen::Vector3f(map->normal[3 * g], map->normal[3 * g + 1], map->normal[3 * g + 2]);
    p.use_raw_distances = true;
    p.mfMinDistance = map->min_dist[g];
    p.mfMaxDistance = map->max_dist[g];
    p.observations = map->has_obs[g] ? 1 : 0;
    p.descriptor = rows32(map->desc + g * 32, 1).clone();
    p.bad = bad && bad[i];
    p.mbTrackInView = track_in_view[i] != 0;
    p.mTrackProjX = proj_x[i]; p.mTrackProjY = proj_y[i]; p.mTrackProjXR = proj_xr[i];
    p.mnTrackScaleLevel = level[i]; p.mTrackViewCos = view_cos[i]; p.mTrackDepth = depth[i];
    p.visible = visible[i];
    p.mnLastFrameSeen = (long unsigned int)last_seen[i];
    T.mvpLocalMapPoints[i] = &p;
  }
  MapPoint outside;
  outside.observations = 1;
  outside.mnId = 999;
  F.mvpMapPoints.assign(F.N, nullptr);
  for (int i = 0; i < F.N; i++)
    if (held && held[i] != -1) F.mvpMapPoints[i] = held[i] >= 0 ? &pts[held[i]] : &outside;
  T.SearchLocalPoints();
  for (int i = 0; i < F.N; i++) {
    const MapPoint* p = F.mvpMapPoints[i];
    assign[i] = !p ? -1 : (p == &outside ? -2 : (int)(p - pts.data()));
  }
  for (int i = 0; i < M; i++) {
    const MapPoint& p = pts[i];
    track_in_view[i] = p.mbTrackInView ? 1 : 0;
    proj_x[i] = p.mTrackProjX; proj_y[i] = p.mTrackProjY; proj_xr[i] = p.mTrackProjXR;
    level[i] = p.mnTrackScaleLevel; view_cos[i] = p.mTrackViewCos; depth[i] = p.mTrackDepth;
    visible[i] = p.visible;
    last_seen[i] = (int32_t)p.mnLastFrameSeen;
    auto it = F.mmProjectPoints.find(p.mnId);
    project_points[2 * i] = it == F.mmProjectPoints.end() ? NAN : it->second.x;
    project_points[2 * i + 1] = it == F.mmProjectPoints.end() ? NAN : it->second.y;
  }
  return (int)F.mmProjectPoints.size();
}

// BASELINE.json configs[3] on the CPU with the reference's own code end to end: both ORBextractor::operator() calls and
// Frame::ComputeStereoMatches (as orbrefsrc_stereo_frame), then Tracking::SearchLocalPoints' work on the left frame
// (src/Tracking.cc:3288-3330): Frame::AssignFeaturesToGrid, Frame::isInFrustum for every point of local map `map_index`
// and ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints). The MapPoint objects of
// a map are built once per thread and reused (in the reference they live in the Atlas), so that the timed work is the
// reference's per-frame work only. Outputs as orbm_stereo_track_frames_batch for one pair.
namespace {
struct MapWorld {
  std::vector<MapPoint> pts;
  std::vector<MapPoint*> ptrs;
  std::vector<uint8_t> skip;
};
MapWorld& map_world(const orbx_local_map* maps, int map_index) {
  thread_local std::map<std::pair<const void*, int>, MapWorld> cache;
  MapWorld& W = cache[{maps->pos, map_index}];
  if ((int)W.pts.size() == maps->m && maps->m > 0) return W;
  const size_t base = (size_t)map_index * maps->m;
  W.pts.assign(maps->m, MapPoint());
  W.ptrs.resize(maps->m);
  W.skip.assign(maps->m, 0);
  for (int i = 0; i < maps->m; i++) {
    const size_t g = base + i;
    MapPoint& p = W.pts[i];
    p.pos = Eigen::Vector3f(maps->pos[3 * g], maps->pos[3 * g + 1], maps->pos[3 * g + 2]);
    p.normal = Eigen::Vector3f(maps->normal[3 * g], maps->normal[3 * g + 1], maps->normal[3 * g + 2]);
    p.use_raw_distances = true;
    p.mfMinDistance = maps->min_dist[g];
    p.mfMaxDistance = maps->max_dist[g];
    p.observations = maps->has_obs[g] ? 1 : 0;
    p.descriptor = rows32(maps->desc + g * 32, 1).clone();
    W.skip[i] = maps->skip ? maps->skip[g] : 0;
    W.ptrs[i] = &p;
  }
  return W;
}
}  // namespace

extern "C" int orbrefsrc_stereo_track_frame(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                                            const unsigned char* img_l, const unsigned char* img_r, int w, int h,
                                            int stride, float mbf, float mb, const orbx_frustum* fr,
                                            const orbx_local_map* maps, int map_index, const uint8_t* occupied,
                                            const orbx_track_params* prm, void* kps_l, unsigned char* desc_l, int* n_l,
                                            void* kps_r, unsigned char* desc_r, int* n_r, float* u_right, float* depth,
                                            int cap, int32_t* assign, int* nmatches, int* n_in_view) {
  ORBextractor left(nfeatures, scale_factor, nlevels, ini_th, min_th), right(nfeatures, scale_factor, nlevels, ini_th, min_th);
  Frame F;
  PinholeStandIn cam;
  F.mpORBextractorLeft = &left;
  F.mpORBextractorRight = &right;
  std::vector<int> lapping = {0, 0};
  cv::Mat mask;
  cv::Mat il(h, w, CV_8UC1, const_cast<unsigned char*>(img_l), (size_t)stride);
  cv::Mat ir(h, w, CV_8UC1, const_cast<unsigned char*>(img_r), (size_t)stride);
  left(il, mask, F.mvKeys, F.mDescriptors, lapping);
  right(ir, mask, F.mvKeysRight, F.mDescriptorsRight, lapping);
  F.N = (int)F.mvKeys.size();
  F.Nleft = -1;
  F.mvScaleFactors = left.GetScaleFactors();
  F.mvInvScaleFactors = left.GetInverseScaleFactors();
  F.mbf = mbf;
  F.mb = mb;
  *n_l = F.N;
  *n_r = (int)F.mvKeysRight.size();
  if (*n_l > cap || *n_r > cap) return -1000;
  F.ComputeStereoMatches();
  // the rest of the Frame constructor that the search reads: mvKeysUn (undistorted camera: a copy, src/Frame.cc:562-571),
  // the grid (:235-236), the pose
  F.mvKeysUn = F.mvKeys;
  Frame::mnMinX = fr->min_x; Frame::mnMaxX = fr->max_x; Frame::mnMinY = fr->min_y; Frame::mnMaxY = fr->max_y;
  Frame::mfGridElementWidthInv = prm->inv_w;
  Frame::mfGridElementHeightInv = prm->inv_h;
  F.AssignFeaturesToGrid();
  cam.mvParameters[0] = fr->fx; cam.mvParameters[1] = fr->fy; cam.mvParameters[2] = fr->cx; cam.mvParameters[3] = fr->cy;
  F.mpCamera = &cam;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) F.mRcw(i, j) = fr->Rcw[3 * i + j];
    F.mtcw(i) = fr->tcw[i];
    F.mOw(i) = fr->Ow[i];
  }
  F.mfLogScaleFactor = fr->log_scale_factor;
  F.mnScaleLevels = fr->n_levels;
  MapPoint occupied_point;
  occupied_point.observations = 1;
  F.mvpMapPoints.assign(F.N, nullptr);
  if (occupied)
    for (int i = 0; i < F.N; i++)
      if (occupied[i]) F.mvpMapPoints[i] = &occupied_point;
  MapWorld& W = map_world(maps, map_index);
  int nToMatch = 0;
  for (int i = 0; i < maps->m; i++) {  // src/Tracking.cc:3288-3300
    MapPoint* pMP = W.ptrs[i];
    pMP->mbTrackInView = false;
    if (W.skip[i]) continue;
    if (F.isInFrustum(pMP, prm->viewing_cos_limit)) nToMatch++;
  }
  *n_in_view = nToMatch;
  *nmatches = 0;
  if (nToMatch > 0) {
    ORBmatcher matcher(prm->nnratio);
    *nmatches = matcher.SearchByProjection(F, W.ptrs, prm->th, prm->far_points != 0, prm->th_far);
  }
  for (int i = 0; i < F.N; i++) {
    const MapPoint* p = F.mvpMapPoints[i];
    assign[i] = (p && p != &occupied_point) ? (int)(p - W.pts.data()) : -1;
  }
  if (*n_l) memcpy(kps_l, F.mvKeys.data(), (size_t)*n_l * sizeof(cv::KeyPoint));
  if (*n_r) memcpy(kps_r, F.mvKeysRight.data(), (size_t)*n_r * sizeof(cv::KeyPoint));
  for (int i = 0; i < *n_l; i++) memcpy(desc_l + (size_t)i * 32, F.mDescriptors.ptr(i), 32);
  for (int i = 0; i < *n_r; i++) memcpy(desc_r + (size_t)i * 32, F.mDescriptorsRight.ptr(i), 32);
  int matched = 0;
  for (int i = 0; i < *n_l; i++) {
    u_right[i] = F.mvuRight[i];
    depth[i] = F.mvDepth[i];
    matched += F.mvuRight[i] >= 0;
  }
  return matched;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, ...) on a two-camera Frame (Nleft != -1): the
// reference's own method, fisheye branches included (src/ORBmatcher.cc:42-221), on a stand-in Frame whose grids are built
// by the reference's own AssignFeaturesToGrid. Outputs as orbref_search_by_projection_map_fisheye.
extern "C" int orbrefsrc_search_by_projection_map_fisheye(const orbx_fisheye_view* fv, const orbx_mappoints* mps,
                                                          const orbx_mappoints_right* mr, float th, float nnratio,
                                                          int far_points, float th_far, int32_t* assign) {
  Frame F;
  const int NL = fv->n_left, NR = fv->n_right, N = NL + NR;
  F.N = N;
  F.Nleft = NL;
  F.mvKeys = keypoints(fv->kps_left, NL);
  F.mvKeysRight = keypoints(fv->kps_right, NR);
  F.mvKeysUn = F.mvKeys;
  F.mDescriptors = rows32(fv->desc, N);
  F.mvuRight.assign(N, -1.f);
  F.mvScaleFactors.assign(fv->scale_factors, fv->scale_factors + fv->n_levels);
  Frame::mnMinX = fv->grid_left.min_x;
  Frame::mnMinY = fv->grid_left.min_y;
  Frame::mfGridElementWidthInv = fv->grid_left.inv_w;
  Frame::mfGridElementHeightInv = fv->grid_left.inv_h;
  F.AssignFeaturesToGrid();
  F.mvLeftToRightMatch.assign(fv->left_to_right, fv->left_to_right + NL);
  F.mvRightToLeftMatch.assign(fv->right_to_left, fv->right_to_left + NR);
  MapPoint occupied;
  occupied.observations = 1;
  F.mvpMapPoints.assign(N, nullptr);
  for (int i = 0; i < N; i++)
    if (fv->occupied && fv->occupied[i]) F.mvpMapPoints[i] = &occupied;
  GeometricCamera camera;
  F.mpCamera = F.mpCamera2 = &camera;
  std::vector<MapPoint> pts(mps->m);
  std::vector<MapPoint*> ptrs(mps->m);
  for (int i = 0; i < mps->m; i++) {
    MapPoint& p = pts[i];
    p.mbTrackInView = mps->track_in_view[i] != 0;
    p.mbTrackInViewR = mr->track_in_view_r[i] != 0;
    p.mTrackProjX = mps->proj_x[i];
    p.mTrackProjY = mps->proj_y[i];
    p.mTrackProjXR = mr->proj_x_r[i];
    p.mTrackProjYR = mr->proj_y_r[i];
    p.mnTrackScaleLevel = mps->level[i];
    p.mnTrackScaleLevelR = mr->level_r[i];
    p.mTrackViewCos = mps->view_cos[i];
    p.mTrackViewCosR = mr->view_cos_r[i];
    p.mTrackDepth = mps->depth[i];
    p.observations = mps->has_obs[i] ? 1 : 0;
    p.descriptor = rows32(mps->desc + (size_t)i * 32, 1);
    ptrs[i] = &p;
  }
  ORBmatcher matcher(nnratio, true);
  const int n = matcher.SearchByProjection(F, ptrs, th, far_points != 0, th_far);
  for (int i = 0; i < N; i++) {
    const MapPoint* p = F.mvpMapPoints[i];
    assign[i] = (p && p != &occupied) ? (int)(p - pts.data()) : -1;
  }
  return n;
}
