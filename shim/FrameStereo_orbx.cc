// shim/FrameStereo_orbx.cc — body of Frame::ComputeStereoMatches (src/Frame.cc:921-1084) forwarding to orbm_stereo_match.
//
// COMPILES ONLY INSIDE THE REFERENCE TREE (needs include/Frame.h and its OpenCV / Eigen / Sophus dependencies, none of
// which exist in this repository's build image). Replace the reference's function body with this file's; everything
// else in Frame.cc stays. What crosses the ABI: mvKeys / mvKeysRight (std::vector<cv::KeyPoint> = 28-byte PODs),
// mDescriptors / mDescriptorsRight (continuous N x 32 CV_8U), mbf, mb; results land in mvuRight / mvDepth.
// The pyramids are NOT copied: orbm_stereo_match reads the device-resident levels of the two extractor handles
// (the reference reads mpORBextractorLeft/Right->mvImagePyramid, :927,1011,1029), so the host mirror can be switched
// off (ORBextractor::SetPyramidMirror(false)).
#include "Frame.h"
#include "orbm.h"
#include "orbx_thread_matcher.h"

namespace ORB_SLAM3 {

void Frame::ComputeStereoMatches() {
  mvuRight = std::vector<float>(N, -1.0f);  // :922-923
  mvDepth = std::vector<float>(N, -1.0f);
  if (N == 0) return;
  int32_t n_matched = 0;
  const int rc = orbm_stereo_match(
      OrbxThreadMatcher(), mpORBextractorLeft->Handle(), mpORBextractorRight->Handle(), /*frame=*/0,
      reinterpret_cast<const orbx_kp*>(mvKeys.data()), mDescriptors.data, N,
      reinterpret_cast<const orbx_kp*>(mvKeysRight.data()), mDescriptorsRight.data, (int)mvKeysRight.size(), mbf, mb,
      mvuRight.data(), mvDepth.data(), &n_matched);
  if (rc != ORBX_OK) throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
}

}  // namespace ORB_SLAM3
