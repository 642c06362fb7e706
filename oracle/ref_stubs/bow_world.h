// Stand-in world for Frame::ComputeBoW (src/Frame.cc:846-851) and KeyFrame::ComputeBoW (src/KeyFrame.cc:98-107): the
// members those two functions touch as plain data, over DBoW2's OWN TemplatedVocabulary (the reference's
// include/ORBVocabulary.h and Thirdparty/DBoW2 are used as they are). Frame.h / KeyFrame.h / Converter.h cannot be used
// here (Eigen, Sophus, g2o, boost), so their include guards are pre-defined. The text of the two functions is cut out of
// the reference files by signature and piped to the compiler (oracle/Makefile); shim/FrameBoW_orbx.cc holds their
// drop-in bodies. TEST INFRASTRUCTURE.
#ifndef ORBREF_STUB_BOW_WORLD_H_
#define ORBREF_STUB_BOW_WORLD_H_
#define FRAME_H
#define KEYFRAME_H
#define CONVERTER_H

#include <vector>

#include <opencv2/opencv.hpp>

#include "ORBVocabulary.h"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

namespace ORB_SLAM3 {
// Converter::toDescriptorVector (src/Converter.cc:26-33): one row view per descriptor
class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors) {
    std::vector<cv::Mat> v;
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
  }
};
struct BowHolder {
  ORBVocabulary* mpORBvocabulary = nullptr;
  cv::Mat mDescriptors;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
};
class Frame : public BowHolder {
 public:
  void ComputeBoW();
};
class KeyFrame : public BowHolder {
 public:
  void ComputeBoW();
};
}  // namespace ORB_SLAM3
using namespace std;  // the reference's headers leak it
#endif
