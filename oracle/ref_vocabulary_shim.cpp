// C entry points over DBoW2's own TemplatedVocabulary<FORB::TDescriptor, FORB> as vendored by the reference
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, FORB.cpp, ...), compiled where it lies against the stand-in OpenCV
// header of oracle/ref_stubs. The tree is loaded with the reference's loadFromTextFile; the transform that runs is
// DBoW2's (the call Frame::ComputeBoW makes, src/Frame.cc:846-851). TEST INFRASTRUCTURE.
#include <string>
#include <vector>

#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Vocabulary;

extern "C" {

void* orbrefsrc_voc_load(const char* path) {
  Vocabulary* v = new Vocabulary();
  if (!v->loadFromTextFile(path)) {
    delete v;
    return nullptr;
  }
  return v;
}
void orbrefsrc_voc_destroy(void* h) { delete static_cast<Vocabulary*>(h); }
int orbrefsrc_voc_size(void* h) { return (int)static_cast<Vocabulary*>(h)->size(); }

// per feature: word id and weight of the leaf reached (transform(feature), getWordWeight), and the node id under which
// transform(features, BowVector, FeatureVector, levelsup) filed the feature, 0xffffffff when it filed it nowhere
// (weight 0: a stopped word, TemplatedVocabulary.h:1154). bow_words / bow_values (cap entries) receive the BowVector.
int orbrefsrc_voc_transform(void* h, const unsigned char* desc, int n, int levelsup, unsigned* word_id, double* weight,
                            unsigned* node_id, unsigned* bow_words, double* bow_values, int cap) {
  Vocabulary* voc = static_cast<Vocabulary*>(h);
  std::vector<cv::Mat> features(n);
  for (int i = 0; i < n; i++) features[i] = cv::Mat(1, 32, CV_8U, const_cast<unsigned char*>(desc) + (size_t)i * 32, 32).clone();
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  voc->transform(features, bv, fv, levelsup);
  for (int i = 0; i < n; i++) {
    word_id[i] = voc->transform(features[i]);
    weight[i] = voc->getWordWeight(word_id[i]);
    node_id[i] = 0xffffffffu;
  }
  for (const auto& kv : fv)
    for (unsigned idx : kv.second) node_id[idx] = kv.first;
  int k = 0;
  for (const auto& kv : bv) {
    if (k < cap) {
      bow_words[k] = kv.first;
      bow_values[k] = kv.second;
    }
    k++;
  }
  return k;
}

#ifdef ORBREF_BOW_WORLD
// Frame::ComputeBoW() / KeyFrame::ComputeBoW() on the stand-in objects of ref_stubs/bow_world.h: in
// liborbref_dbow2_src.so the two functions are the reference's own text (src/Frame.cc:846-851, src/KeyFrame.cc:98-107,
// piped in), in the shim bow worlds they are the drop-in bodies of shim/FrameBoW_orbx.cc. which: 0 Frame, 1 KeyFrame,
// 2 KeyFrame whose mBowVec is already filled but whose mFeatVec is empty (recomputed, :100), 3 Frame whose mBowVec is
// already filled (left alone, :847). Outputs: the BowVector as (bow_words, bow_values)[cap], the FeatureVector as
// node_id[n] per feature (0xffffffff = filed nowhere). Returns the BowVector's size.
int orbrefsrc_compute_bow(void* h, const unsigned char* desc, int n, int which, unsigned* node_id, unsigned* bow_words,
                          double* bow_values, int cap) {
  ORB_SLAM3::Frame F;
  ORB_SLAM3::KeyFrame K;
  ORB_SLAM3::BowHolder& B = (which == 0 || which == 3) ? static_cast<ORB_SLAM3::BowHolder&>(F) : K;
  B.mpORBvocabulary = static_cast<Vocabulary*>(h);
  B.mDescriptors = cv::Mat(n, 32, CV_8U, const_cast<unsigned char*>(desc), 32).clone();
  if (which >= 2) B.mBowVec.addWeight(7, 0.25);
  if (which == 0 || which == 3)
    F.ComputeBoW();
  else
    K.ComputeBoW();
  for (int i = 0; i < n; i++) node_id[i] = 0xffffffffu;
  for (const auto& kv : B.mFeatVec)
    for (unsigned idx : kv.second) node_id[idx] = kv.first;
  int k = 0;
  for (const auto& kv : B.mBowVec) {
    if (k < cap) {
      bow_words[k] = kv.first;
      bow_values[k] = kv.second;
    }
    k++;
  }
  return k;
}
#endif
}
