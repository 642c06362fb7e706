// shim/FrameStereo_orbx.cc — bodies of Frame::ComputeStereoMatches (src/Frame.cc:921-1084) forwarding to
// orbm_stereo_match, and of Frame::ComputeStereoFishEyeMatches (:1271-1331) forwarding to orbm_knn2.
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

// Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1271-1331): cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2) over the
// lapping-area descriptors -> orbm_knn2 (POPC kernel at these sizes, the tensor-core form above ~6k x 6k); Lowe's ratio,
// the triangulation call (the reference's own KannalaBrandt8 object — camera models are not rebuilt) and the
// bookkeeping stay here, in the reference's statement order.
void Frame::ComputeStereoFishEyeMatches() {
  mvLeftToRightMatch = std::vector<int>(Nleft, -1);  // :1281-1286
  mvRightToLeftMatch = std::vector<int>(Nright, -1);
  mvDepth = std::vector<float>(Nleft, -1.0f);
  mvuRight = std::vector<float>(Nleft, -1);
  mvStereo3Dpoints = std::vector<Eigen::Vector3f>(Nleft);
  mnCloseMPs = 0;
  const int nq = mDescriptors.rows - monoLeft, nt = mDescriptorsRight.rows - monoRight;  // :1276-1278
  if (nq <= 0) return;
  std::vector<int32_t> i1(nq), d1(nq), i2(nq), d2(nq);
  if (orbm_knn2(OrbxThreadMatcher(), mDescriptors.ptr(monoLeft), nq, nt > 0 ? mDescriptorsRight.ptr(monoRight) : nullptr,
                nt > 0 ? nt : 0, i1.data(), d1.data(), i2.data(), d2.data()) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  for (int q = 0; q < nq; q++) {
    // (*it).size() >= 2 && (*it)[0].distance < (*it)[1].distance * 0.7: float distances, the product in double  :1299
    if (i2[q] < 0 || !((float)d1[q] < (float)d2[q] * 0.7)) continue;
    const int il = q + monoLeft, ir = i1[q] + monoRight;
    Eigen::Vector3f p3D;
    const float sigma1 = mvLevelSigma2[mvKeys[il].octave], sigma2 = mvLevelSigma2[mvKeysRight[ir].octave];
    const float depth = static_cast<KannalaBrandt8*>(mpCamera)->TriangulateMatches(mpCamera2, mvKeys[il], mvKeysRight[ir],
                                                                                   mRlr, mtlr, sigma1, sigma2, p3D);
    if (depth > 0.0001f) {  // :1317-1325
      mvLeftToRightMatch[il] = ir;
      mvRightToLeftMatch[ir] = il;
      mvStereo3Dpoints[il] = p3D;
      mvDepth[il] = depth;
    }
  }
}

}  // namespace ORB_SLAM3
