// shim/ORBextractor.h — header-compatible replacement for the reference's include/ORBextractor.h (:48-120).
//
// Same namespace, class name, constructor, operator(), getters and the public mvImagePyramid member, so that
// src/Frame.cc, src/Tracking.cc and every other caller compile and link unchanged; the work is forwarded to the orbx C
// ABI (include/orbx.h, liborbx.so, hand-written sm_100a CUDA). Nothing is computed on the host: if no CUDA device is
// present the constructor throws std::runtime_error with the ABI's error text.
//
// Differences a maintainer should know about (INTEGRATION.md has the details):
//  * the protected helpers of the reference (ComputePyramid, ComputeKeyPointsOctTree, DistributeOctTree, ExtractorNode)
//    do not exist — nothing outside ORBextractor.cc uses them;
//  * mvImagePyramid is a host MIRROR that is downloaded after every call (0.95 MB at 640x480) because
//    Frame::ComputeStereoMatches reads it (src/Frame.cc:927,1011,1024,1029). When ComputeStereoMatches is replaced by
//    orbm_stereo_match (shim/FrameStereo_orbx.cc) the mirror is dead weight: SetPyramidMirror(false) switches it off.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <opencv2/opencv.hpp>

#include <vector>

struct orbx_extractor;

namespace ORB_SLAM3 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor();
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // Compute the ORB features and descriptors on an image. Mask is ignored, as in the reference.
  // Returns monoIndex (src/ORBextractor.cc:1105), or -1 for an empty image (:1021).
  int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                 cv::OutputArray _descriptors, std::vector<int>& vLappingArea);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return (float)scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  std::vector<cv::Mat> mvImagePyramid;  // ROI at (19, 19) of a (w + 38) x (h + 38) buffer, like :1114-1118

  // ---- additions (not in the reference) ----
  void SetPyramidMirror(bool on) { mbMirrorPyramid = on; }
  orbx_extractor* Handle() const { return mpHandle; }  // for orbm_stereo_match (device-resident pyramid)
  static void SetDevice(int cuda_ordinal);             // device used by extractors constructed afterwards (default 0)
  static int GetDevice();                              // ... and by the per-thread matcher contexts (orbx_thread_matcher.h)

 protected:
  int nfeatures;
  double scaleFactor;  // a double initialised from a float, as in the reference (include/ORBextractor.h:106)
  int nlevels;
  int iniThFAST;
  int minThFAST;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<float> mvScaleFactor;
  std::vector<float> mvInvScaleFactor;
  std::vector<float> mvLevelSigma2;
  std::vector<float> mvInvLevelSigma2;

 private:
  orbx_extractor* mpHandle;
  bool mbMirrorPyramid;
  std::vector<cv::Mat> mvBordered;  // owners of the mirror's pixel memory
  cv::Mat mContinuous;              // scratch when the input is not a plain 8-bit single-channel matrix
};

}  // namespace ORB_SLAM3

#endif
