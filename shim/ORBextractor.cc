// shim/ORBextractor.cc — ORB_SLAM3::ORBextractor forwarding to the orbx C ABI. Replaces the reference's
// src/ORBextractor.cc in libORB_SLAM3.so (CMakeLists.txt:69-75); link with -lorbx.
#include "ORBextractor.h"

#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "orbx.h"

namespace ORB_SLAM3 {

namespace {
std::atomic<int> g_device{0};
const int EDGE_THRESHOLD = 19;  // src/ORBextractor.cc:73
}  // namespace

void ORBextractor::SetDevice(int cuda_ordinal) { g_device = cuda_ordinal; }
int ORBextractor::GetDevice() { return g_device; }

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
      minThFAST(_minThFAST), mpHandle(nullptr), mbMirrorPyramid(std::getenv("ORBX_SHIM_NO_PYRAMID_MIRROR") == nullptr) {
  const int rc = orbx_extractor_create(&mpHandle, g_device, nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST, 1);
  if (rc != ORBX_OK)
    throw std::runtime_error(std::string("ORBextractor: orbx_extractor_create failed: ") + orbx_last_error(nullptr));
  mvScaleFactor.resize(nlevels);
  mvInvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels);
  mvInvLevelSigma2.resize(nlevels);
  mnFeaturesPerLevel.resize(nlevels);
  if (orbx_extractor_tables(mpHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                            mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()) != ORBX_OK)
    throw std::runtime_error("ORBextractor: orbx_extractor_tables failed");
  mvImagePyramid.resize(nlevels);  // :416
  mvBordered.resize(nlevels);
}

ORBextractor::~ORBextractor() { orbx_extractor_destroy(mpHandle); }

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
  if (_image.empty()) return -1;  // :1021
  cv::Mat image = _image.getMat();
  // the reference asserts CV_8UC1 (:1024); a non-continuous-row matrix is fine (stride is passed through)
  if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");
  const int lap0 = vLappingArea.size() > 0 ? vLappingArea[0] : 0;
  const int lap1 = vLappingArea.size() > 1 ? vLappingArea[1] : 0;

  const int cap = orbx_extractor_capacity(mpHandle);
  static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_kp), "cv::KeyPoint must be the 28-byte POD the ABI writes");
  _keypoints.resize(cap);
  cv::Mat desc(cap, 32, CV_8U);
  int32_t n = 0, mono = 0;
  const int rc = orbx_extract(mpHandle, image.data, image.cols, image.rows, (int)image.step, lap0, lap1,
                              reinterpret_cast<orbx_kp*>(_keypoints.data()), desc.data, cap, &n, &mono);
  if (rc == ORBX_E_EMPTY) {
    _keypoints.clear();
    return -1;
  }
  if (rc != ORBX_OK) throw std::runtime_error(std::string("ORBextractor: ") + orbx_last_error(mpHandle));
  _keypoints.resize(n);                 // _keypoints = vector<KeyPoint>(nkeypoints)            :1059
  if (n == 0) {
    _descriptors.release();             //                                                       :1050-1052
  } else {
    _descriptors.create(n, 32, CV_8U);  //                                                       :1053-1055
    cv::Mat out = _descriptors.getMat();
    for (int i = 0; i < n; i++) memcpy(out.ptr(i), desc.ptr(i), 32);
  }

  if (mbMirrorPyramid) {
    for (int level = 0; level < nlevels; ++level) {
      int w = 0, h = 0;
      orbx_level_size(mpHandle, level, &w, &h);
      cv::Mat& whole = mvBordered[level];
      whole.create(h + 2 * EDGE_THRESHOLD, w + 2 * EDGE_THRESHOLD, CV_8UC1);
      orbx_download_pyramid(mpHandle, 0, level, whole.data, (int)whole.step);
      mvImagePyramid[level] = whole(cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));  // :1116-1118
    }
  }
  return mono;
}

}  // namespace ORB_SLAM3
