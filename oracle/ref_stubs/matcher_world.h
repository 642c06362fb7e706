// Stand-in world for compiling the reference's src/ORBmatcher.cc where it lies (oracle/Makefile, target `ref`): the
// headers Frame.h / KeyFrame.h / MapPoint.h pull in Eigen, Sophus, DBoW2's vocabulary, g2o and boost and cannot be used
// here, so their include guards are pre-defined and the classes below offer exactly the members ORBmatcher.cc touches,
// as plain data the test shim fills in (their grid functions are the reference's own text, piped in by the Makefile). The matcher code that runs against them is the reference's own. TEST
// INFRASTRUCTURE. Geometry types are small value types; the tests drive the projecting overloads with identity poses
// and an orthographic stand-in camera so that no result depends on how a 3x3 product is associated.
#ifndef ORBREF_STUB_MATCHER_WORLD_H_
#define ORBREF_STUB_MATCHER_WORLD_H_
#define FRAME_H
#define KEYFRAME_H
#define MAPPOINT_H
#define TRACKING_H
#define ATLAS_H
#define LOCALMAPPING_H
#define SYSTEM_H
#define MAP_H
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include <opencv2/opencv.hpp>
#include <tbb/parallel_for.h>

#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "ORBextractor.h"  // the reference's own header: Frame::ComputeStereoMatches reads mvImagePyramid of both

namespace Eigen {
struct Vector2f {
  float v[2];
  Vector2f() : v{0, 0} {}
  Vector2f(float a, float b) : v{a, b} {}
  float operator()(int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
};
struct Vector3f {
  float v[3];
  Vector3f() : v{0, 0, 0} {}
  Vector3f(float a, float b, float c) : v{a, b, c} {}
  float operator()(int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
  Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  Vector3f operator+(const Vector3f& o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  Vector3f operator/(float s) const { return Vector3f(v[0] / s, v[1] / s, v[2] / s); }
  Vector3f operator*(float s) const { return Vector3f(v[0] * s, v[1] * s, v[2] * s); }
  Vector3f operator-() const { return Vector3f(-v[0], -v[1], -v[2]); }
  // Eigen reduces a fixed-size 3-vector without packets as c0 + (c1 + c2) (redux_novec_unroller halves the range,
  // Eigen/src/Core/Redux.h); Frame::isInFrustum's results depend on it (the other tests use identity poses)
  float dot(const Vector3f& o) const { return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); }
  float norm() const { return std::sqrt(dot(*this)); }
};
struct Matrix3f {
  float m[3][3];
  Matrix3f() : m{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}} {}
  float operator()(int i, int j) const { return m[i][j]; }
  float& operator()(int i, int j) { return m[i][j]; }
  Vector3f operator*(const Vector3f& p) const {
    return Vector3f(m[0][0] * p.v[0] + (m[0][1] * p.v[1] + m[0][2] * p.v[2]),
                    m[1][0] * p.v[0] + (m[1][1] * p.v[1] + m[1][2] * p.v[2]),
                    m[2][0] * p.v[0] + (m[2][1] * p.v[1] + m[2][2] * p.v[2]));
  }
  Matrix3f operator*(const Matrix3f& o) const {
    Matrix3f r;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j];
    return r;
  }
  // `m << a, b, c, ...;` (row major) and inverse(): used by the reference-side shim bodies (shim/*.cc), not by the matcher
  struct CommaInit {
    Matrix3f* m;
    int at;
    CommaInit& operator,(float x) {
      m->m[at / 3][at % 3] = x;
      at++;
      return *this;
    }
  };
  CommaInit operator<<(float x) {
    m[0][0] = x;
    return CommaInit{this, 1};
  }
  Matrix3f inverse() const {
    const float a = m[0][0], b = m[0][1], c = m[0][2], d = m[1][0], e = m[1][1], f = m[1][2], g = m[2][0], h = m[2][1],
                i = m[2][2];
    const float det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    Matrix3f r;
    r.m[0][0] = (e * i - f * h) / det; r.m[0][1] = (c * h - b * i) / det; r.m[0][2] = (b * f - c * e) / det;
    r.m[1][0] = (f * g - d * i) / det; r.m[1][1] = (a * i - c * g) / det; r.m[1][2] = (c * d - a * f) / det;
    r.m[2][0] = (d * h - e * g) / det; r.m[2][1] = (b * g - a * h) / det; r.m[2][2] = (a * e - b * d) / det;
    return r;
  }
  Matrix3f transpose() const {
    Matrix3f r;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r.m[i][j] = m[j][i];
    return r;
  }
};
// Eigen::Matrix<float, 3, 1> / <float, 3, 3> as Frame::isInFrustum spells them
template <int R, int C> struct MatSel;
template <> struct MatSel<3, 1> { typedef Vector3f type; };
template <> struct MatSel<3, 3> { typedef Matrix3f type; };
template <typename T, int R, int C> using Matrix = typename MatSel<R, C>::type;
}  // namespace Eigen

namespace Sophus {
struct SE3f {
  Eigen::Matrix3f R;
  Eigen::Vector3f t;
  SE3f() {}
  SE3f(const Eigen::Matrix3f& R_, const Eigen::Vector3f& t_) : R(R_), t(t_) {}
  Eigen::Matrix3f rotationMatrix() const { return R; }
  Eigen::Vector3f translation() const { return t; }
  SE3f inverse() const { return SE3f(R.transpose(), -(R.transpose() * t)); }
  Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return R * p + t; }
  SE3f operator*(const SE3f& o) const { return SE3f(R * o.R, R * o.t + t); }
};
template <typename T>
struct Sim3 {
  Eigen::Matrix3f R;
  Eigen::Vector3f t;
  float s = 1.f;
  Eigen::Matrix3f rotationMatrix() const { return R; }
  Eigen::Vector3f translation() const { return t; }
  float scale() const { return s; }
  Sim3 inverse() const {
    Sim3 r;
    r.R = R.transpose();
    r.s = 1.f / s;
    r.t = -((R.transpose() * t) * (1.f / s));
    return r;
  }
  Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return (R * p) * s + t; }
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus

namespace ORB_SLAM3 {
class Frame;
class KeyFrame;
class MapPoint;

class GeometricCamera {
 public:
  virtual ~GeometricCamera() {}
  // orthographic stand-in: the image point of (x, y, z) is (x, y); see the header comment
  virtual Eigen::Vector2f project(const Eigen::Vector3f& p) { return Eigen::Vector2f(p(0), p(1)); }
  virtual Eigen::Matrix3f toK_() { return Eigen::Matrix3f(); }
  // GeometricCamera::getParameter (include/CameraModels/GeometricCamera.h:81): fx, fy, cx, cy of a pinhole model
  virtual float getParameter(const int i) { return i < 2 ? 1.f : 0.f; }
  // Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:122-149) with F12 supplied by the test
  float F12[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  // two-camera rigs: `tag` says which camera of its KeyFrame this is; against a second camera with tag 1 the test uses
  // F12b, so a test can tell which of the four (left / right) x (left / right) pairs the matcher selected (:1007-1043)
  int tag = 0;
  float F12b[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  virtual bool epipolarConstrain(GeometricCamera* other, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2,
                                 const Eigen::Matrix3f&, const Eigen::Vector3f&, const float, const float unc) {
    const float* F12 = (other && other->tag) ? this->F12b : this->F12;
    const float a = kp1.pt.x * F12[0] + kp1.pt.y * F12[3] + F12[6];
    const float b = kp1.pt.x * F12[1] + kp1.pt.y * F12[4] + F12[7];
    const float c = kp1.pt.x * F12[2] + kp1.pt.y * F12[5] + F12[8];
    const float num = a * kp2.pt.x + b * kp2.pt.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * unc;
  }
};

// Pinhole::project(const Eigen::Vector3f&) (src/CameraModels/Pinhole.cpp:47-53) with the reference's expression; the
// reference's Pinhole.cpp needs GeometricCamera.h (Eigen, boost serialization) and cannot be compiled here
class PinholeStandIn : public GeometricCamera {
 public:
  float mvParameters[4] = {1, 1, 0, 0};
  float getParameter(const int i) override { return mvParameters[i]; }
  Eigen::Vector2f project(const Eigen::Vector3f& v3D) override {
    Eigen::Vector2f res;
    res(0) = mvParameters[0] * v3D(0) / v3D(2) + mvParameters[2];
    res(1) = mvParameters[1] * v3D(1) / v3D(2) + mvParameters[3];
    return res;
  }
};

// KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp) is camera-model code (out of scope, and it
// needs Eigen's SVD): the stand-in returns a deterministic pseudo-depth of its inputs — positive for about two thirds of
// the pairs — and a p3D that encodes both keypoints, so that Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1271-1331)
// can be checked for WHICH pairs it triangulates, with which sigmas, and where it files the answers.
class KannalaBrandt8 : public GeometricCamera {
 public:
  float TriangulateMatches(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2,
                           const Eigen::Matrix3f& R12, const Eigen::Vector3f& t12, const float sigmaLevel, const float unc,
                           Eigen::Vector3f& p3D) {
    const float d = kp1.pt.x - kp2.pt.x + 0.25f * (kp1.pt.y - kp2.pt.y) + t12(0);
    p3D = Eigen::Vector3f(kp1.pt.x * sigmaLevel, kp2.pt.y * unc, d * R12(0, 0));
    return pCamera2 ? std::fmod(std::fabs(d), 3.0f) - 1.0f : -1.0f;
  }
};

// std::mutex members would make the stand-ins immovable; the tests are single threaded
struct CopyableMutex : std::mutex {
  CopyableMutex() {}
  CopyableMutex(const CopyableMutex&) {}
  CopyableMutex& operator=(const CopyableMutex&) { return *this; }
};

class MapPoint {
 public:
  // MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:372-441) is the reference's own text, piped in at build time
  CopyableMutex mMutexFeatures;
  bool mbBad = false;
  std::map<KeyFrame*, std::tuple<int, int>> mObservations;
  cv::Mat mDescriptor;
  void ComputeDistinctiveDescriptors();
  // what Frame::isInFrustum leaves on the point (include/MapPoint.h:172-180)
  bool mbTrackInView = false, mbTrackInViewR = false;
  float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackProjYR = 0, mTrackDepth = 0, mTrackDepthR = 0;
  int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
  float mTrackViewCos = 0, mTrackViewCosR = 0;
  long unsigned int mnLastFrameSeen = 0, mnFuseCandidateForKF = 0;
  // state behind the accessors
  bool bad = false;
  int observations = 0, predicted_level = 0;
  cv::Mat descriptor;
  Eigen::Vector3f pos, normal;
  float min_dist = 0, max_dist = 1e30f;
  std::set<const KeyFrame*> in_keyframes;
  int added_to = -1;          // AddObservation(pKF, idx) / Replace() records, for the Fuse tests
  MapPoint* replaced_by = nullptr;

  bool isBad() { return bad; }
  int Observations() { return observations; }
  cv::Mat GetDescriptor() { return descriptor.clone(); }
  Eigen::Vector3f GetWorldPos() { return pos; }
  Eigen::Vector3f GetNormal() { return normal; }
  bool use_raw_distances = false;  // isInFrustum test: the reference's accessors (src/MapPoint.cc:533-541)
  float GetMinDistanceInvariance() { return use_raw_distances ? 0.8f * mfMinDistance : min_dist; }
  float GetMaxDistanceInvariance() { return use_raw_distances ? 1.2f * mfMaxDistance : max_dist; }
  int PredictScale(const float&, KeyFrame*) { return predicted_level; }
  int PredictScale(const float&, Frame*) { return predicted_level; }
  // MapPoint::PredictScale(const float&, Frame*) (src/MapPoint.cc:559-573) as the reference's own text, piped in under
  // this name (the matcher tests above fix the predicted level instead); isInFrustum's call is renamed with it
  int PredictScaleRef(const float& currentDist, Frame* pF);
  CopyableMutex mMutexPos;
  float mfMaxDistance = 0, mfMinDistance = 0;
  bool IsInKeyFrame(KeyFrame* kf) { return in_keyframes.count(kf) != 0; }
  std::map<const KeyFrame*, int> index_in;
  std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* kf) {
    auto it = index_in.find(kf);
    return std::make_tuple(it == index_in.end() ? -1 : it->second, -1);
  }
  void AddObservation(KeyFrame*, int idx) { added_to = idx; }
  // what Tracking::SearchLocalPoints (src/Tracking.cc:3249-3330) touches besides the tracking fields above
  long unsigned int mnId = 0;
  int visible = 0;  // mnVisible
  void IncreaseVisible(int n = 1) { visible += n; }
  // the two accessors shim/Tracking_orbx.cc asks a maintainer to add next to GetMinDistanceInvariance (mfMinDistance /
  // mfMaxDistance are protected, include/MapPoint.h:241-242; PredictScale reads the raw value, src/MapPoint.cc:559-573)
  float GetMinDistanceRaw() { return mfMinDistance; }
  float GetMaxDistanceRaw() { return mfMaxDistance; }
  void Replace(MapPoint* other) { replaced_by = other; }
};

// members shared by the Frame and KeyFrame stand-ins
class FeatureHolder {
 public:
  int N = 0, Nleft = -1, NLeft = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors;
  std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera* mpCamera = nullptr;
  GeometricCamera* mpCamera2 = nullptr;
  float fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0, mb = 0;
  float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
  Sophus::SE3f pose;
  bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }
  Sophus::SE3f GetPose() const { return pose; }
  Sophus::SE3f GetPoseInverse() const { return pose.inverse(); }
  Sophus::SE3f GetRightPose() const { return pose; }
  Sophus::SE3f GetRightPoseInverse() const { return pose.inverse(); }
  Eigen::Vector3f GetCameraCenter() const { return pose.inverse().translation(); }
  Eigen::Vector3f GetRightCameraCenter() const { return pose.inverse().translation(); }
  Sophus::SE3f GetRelativePoseTrl() const { return Sophus::SE3f(); }
};

class Frame : public FeatureHolder {
 public:
  // what Frame::ComputeStereoMatches (src/Frame.cc:921-1084) reads and writes besides the shared members
  ORBextractor* mpORBextractorLeft = nullptr;
  ORBextractor* mpORBextractorRight = nullptr;
  cv::Mat mDescriptorsRight;
  std::vector<float> mvInvScaleFactors;
  // Frame::isInFrustum (src/Frame.cc:632-699): the reference's own text, piped in at build time
  Eigen::Matrix3f mRcw;
  Eigen::Vector3f mtcw, mOw;
  float mfLogScaleFactor = 0;
  int mnScaleLevels = 0;
  bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
  Eigen::Vector3f GetOw() const { return mOw; }  // include/Frame.h:187
  long unsigned int mnId = 0;
  std::map<long unsigned int, cv::Point2f> mmProjectPoints;  // include/Frame.h:321
  bool isInFrustumChecks(MapPoint*, float, bool = false) { return false; }
  void ComputeStereoMatches();  // defined by the reference's own text, piped in at build time (oracle/Makefile)
  // Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1271-1331): the reference's own text, piped in likewise
  void ComputeStereoFishEyeMatches();
  int Nright = 0, monoLeft = 0, monoRight = 0, mnCloseMPs = 0;
  std::vector<Eigen::Vector3f> mvStereo3Dpoints;
  Eigen::Matrix3f mRlr;
  Eigen::Vector3f mtlr;
  cv::BFMatcher BFmatcher;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
  // the grid (include/Frame.h:279, 372-373; the reference keeps the bounds static): AssignFeaturesToGrid, PosInGrid and
  // GetFeaturesInArea are the reference's own text, piped in at build time (src/Frame.cc:520-547, 833-844, 765-831)
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  // static like the reference's (include/Frame.h:372-378); they hide the per-object fields of the shared base
  static inline float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
  static inline float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  void AssignFeaturesToGrid();
  bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                        const int maxLevel = -1, const bool bRight = false) const;
};

class KeyFrame : public FeatureHolder {
 public:
  std::vector<MapPoint*> mvpMapPoints;
  bool isBad() { return false; }
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  std::set<MapPoint*> GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints)
      if (p) s.insert(p);
    return s;
  }
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  bool record_only = true;  // the Fuse tests read the match off MapPoint::added_to and leave the graph alone
  void AddMapPoint(MapPoint* p, const size_t& idx) {
    if (!record_only) mvpMapPoints[idx] = p;
  }
  // KeyFrame::GetFeaturesInArea is the reference's own text too (src/KeyFrame.cc:705-749); the grid is the Frame's,
  // copied as the KeyFrame constructor does (src/KeyFrame.cc:109-121)
  int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  std::vector<std::vector<std::vector<size_t>>> mGrid, mGridRight;
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
  // the accessor shim/ORBmatcher_next_orbx.cc asks a maintainer to add next to mGrid
  const std::vector<size_t>& GetGridCell(int c, int r) const { return mGrid[c][r]; }
  void take_grid(const Frame& F) {
    mnMinX = F.mnMinX;
    mnMinY = F.mnMinY;
    mfGridElementWidthInv = F.mfGridElementWidthInv;
    mfGridElementHeightInv = F.mfGridElementHeightInv;
    mGrid.assign(mnGridCols, std::vector<std::vector<size_t>>(mnGridRows));
    mGridRight.assign(mnGridCols, std::vector<std::vector<size_t>>(mnGridRows));
    for (int i = 0; i < mnGridCols; i++)
      for (int j = 0; j < mnGridRows; j++) {
        mGrid[i][j] = F.mGrid[i][j];
        mGridRight[i][j] = F.mGridRight[i][j];  // (src/KeyFrame.cc:109-121 copies both)
      }
  }
  // the right camera's cell, for the two-camera branch of shim Fuse (accessor to add next to mGridRight)
  const std::vector<size_t>& GetGridCellRight(int c, int r) const { return mGridRight[c][r]; }
};

// The members Tracking::SearchLocalPoints (src/Tracking.cc:3249-3330) reads, as plain data: the reference's own text of
// that function is piped in at build time (oracle/Makefile); shim/Tracking_orbx.cc is its drop-in body.
class ORBmatcher;
class System {
 public:
  enum eSensor { MONOCULAR = 0, STEREO = 1, RGBD = 2, IMU_MONOCULAR = 3, IMU_STEREO = 4, IMU_RGBD = 5 };  // include/System.h:87-94
};
class Map {
 public:
  bool inertial_ba2 = false;
  bool GetIniertialBA2() { return inertial_ba2; }
};
class Atlas {
 public:
  Map map;
  bool imu_initialized = false;
  Map* GetCurrentMap() { return &map; }
  bool isImuInitialized() { return imu_initialized; }
};
class LocalMapping {
 public:
  bool mbFarPoints = false;
  float mThFarPoints = 0.f;
};
class Tracking {
 public:
  enum eTrackingState {  // include/Tracking.h:122-130
    SYSTEM_NOT_READY = -1, NO_IMAGES_YET = 0, NOT_INITIALIZED = 1, OK = 2, RECENTLY_LOST = 3, LOST = 4, OK_KLT = 5
  };
  eTrackingState mState = OK;
  int mSensor = System::STEREO;
  Frame mCurrentFrame;
  std::vector<MapPoint*> mvpLocalMapPoints;
  Atlas* mpAtlas = nullptr;
  LocalMapping* mpLocalMapper = nullptr;
  unsigned int mnLastRelocFrameId = 0;
  void SearchLocalPoints();
};
}  // namespace ORB_SLAM3
using namespace std;  // the reference's headers leak it; ORBmatcher.h relies on that
#endif
