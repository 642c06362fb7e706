// C entry points over the reference's own ORB_SLAM3::ORBextractor (compiled from /root/reference/src/ORBextractor.cc with
// the stand-in headers of oracle/ref_stubs). TEST INFRASTRUCTURE: used by tests/test_oracle_vs_reference_source.py to
// check the oracle's restatement against the reference code itself. Built only where /root/reference exists.
#include <cstring>
#include <vector>

#include "ORBextractor.h"

namespace {
struct Access : ORB_SLAM3::ORBextractor {
  using ORB_SLAM3::ORBextractor::ORBextractor;
  using ORB_SLAM3::ORBextractor::DistributeOctTree;
  using ORB_SLAM3::ORBextractor::ComputePyramid;
  using ORB_SLAM3::ORBextractor::ComputeKeyPointsOctTree;
  using ORB_SLAM3::ORBextractor::ComputeKeyPointsOctTree_;
  using ORB_SLAM3::ORBextractor::mnFeaturesPerLevel;
  using ORB_SLAM3::ORBextractor::umax;
};
}  // namespace

extern "C" {

void* orbrefsrc_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
  return new Access(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void orbrefsrc_destroy(void* h) { delete static_cast<Access*>(h); }

// scale / inv_scale / sigma2 / inv_sigma2: nlevels floats; per_level: nlevels ints; umax: 16 ints
void orbrefsrc_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* per_level,
                      int* umax) {
  Access* ex = static_cast<Access*>(h);
  const int n = ex->GetLevels();
  const std::vector<float> a = ex->GetScaleFactors(), b = ex->GetInverseScaleFactors(), c = ex->GetScaleSigmaSquares(),
                           d = ex->GetInverseScaleSigmaSquares();
  for (int i = 0; i < n; i++) {
    scale[i] = a[i];
    inv_scale[i] = b[i];
    sigma2[i] = c[i];
    inv_sigma2[i] = d[i];
    per_level[i] = ex->mnFeaturesPerLevel[i];
  }
  for (int i = 0; i < 16; i++) umax[i] = ex->umax[i];
}

// int ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea). Returns the reference's return value;
// *n_out = keypoints.size(); kps = cv::KeyPoint records (28 B), desc = n x 32.
int orbrefsrc_extract(void* h, const unsigned char* img, int w, int h_img, int stride, int lap0, int lap1, void* kps,
                      unsigned char* desc, int cap, int* n_out) {
  Access* ex = static_cast<Access*>(h);
  cv::Mat image(h_img, w, CV_8UC1, const_cast<unsigned char*>(img), (size_t)stride), mask, descriptors;
  if (!img || w <= 0 || h_img <= 0) image = cv::Mat();
  std::vector<cv::KeyPoint> keypoints;
  std::vector<int> lapping = {lap0, lap1};
  const int mono = (*ex)(image, mask, keypoints, descriptors, lapping);
  *n_out = (int)keypoints.size();
  if (*n_out > cap) return -1000;
  if (*n_out > 0) {
    memcpy(kps, keypoints.data(), keypoints.size() * sizeof(cv::KeyPoint));
    for (int i = 0; i < *n_out; i++) memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
  }
  return mono;
}

// std::vector<cv::KeyPoint> ORBextractor::DistributeOctTree(vToDistributeKeys, minX, maxX, minY, maxY, N, level)
int orbrefsrc_distribute(void* h, const void* kps_in, int n_in, int min_x, int max_x, int min_y, int max_y, int n_want,
                         int level, void* kps_out, int cap) {
  Access* ex = static_cast<Access*>(h);
  std::vector<cv::KeyPoint> in(n_in);
  if (n_in) memcpy(static_cast<void*>(in.data()), kps_in, (size_t)n_in * sizeof(cv::KeyPoint));
  const std::vector<cv::KeyPoint> out = ex->DistributeOctTree(in, min_x, max_x, min_y, max_y, n_want, level);
  if ((int)out.size() > cap) return -1000;
  if (!out.empty()) memcpy(kps_out, out.data(), out.size() * sizeof(cv::KeyPoint));
  return (int)out.size();
}

// ComputePyramid + one of the two keypoint stages: serial_twin != 0 runs ComputeKeyPointsOctTree (:886-999), else the
// TBB twin ComputeKeyPointsOctTree_ (:759-885, the one operator() calls) under the serial executor. Keypoints of all
// levels, level by level (level coordinates, octave / size / angle set); counts[nlevels] = keypoints per level.
int orbrefsrc_keypoints(void* h, const unsigned char* img, int w, int h_img, int stride, int serial_twin, void* kps,
                        int cap, int* counts) {
  Access* ex = static_cast<Access*>(h);
  cv::Mat image(h_img, w, CV_8UC1, const_cast<unsigned char*>(img), (size_t)stride);
  ex->ComputePyramid(image);
  std::vector<std::vector<cv::KeyPoint>> all;
  if (serial_twin) ex->ComputeKeyPointsOctTree(all);
  else ex->ComputeKeyPointsOctTree_(all);
  int n = 0;
  for (size_t l = 0; l < all.size(); l++) {
    counts[l] = (int)all[l].size();
    if (n + counts[l] > cap) return -1000;
    if (counts[l]) memcpy(static_cast<char*>(kps) + (size_t)n * sizeof(cv::KeyPoint), all[l].data(), all[l].size() * sizeof(cv::KeyPoint));
    n += counts[l];
  }
  return n;
}
}
