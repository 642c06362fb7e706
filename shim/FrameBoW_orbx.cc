// shim/FrameBoW_orbx.cc — drop-in bodies of Frame::ComputeBoW() (src/Frame.cc:846-851) and KeyFrame::ComputeBoW()
// (src/KeyFrame.cc:98-107): mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) with the per-feature tree
// descent (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1262, six levels of 10-way Hamming comparisons per
// descriptor) on the device — SURVEY.md §8(f) rank 2, the producer of the FeatureVector that SearchByBoW and
// SearchForTriangulation consume.
//
// COMPILES ONLY INSIDE THE REFERENCE TREE (Frame.h, KeyFrame.h, ORBVocabulary.h). Replace the two bodies with the ones
// below; DBoW2 itself is not touched: the tree is read through a derived class (its members are protected, not
// private). The vocabulary is flattened and uploaded once per thread (orbm_set_vocabulary; ORBvoc: 1.1 M nodes, 35 MB of
// descriptors), then every call is one orbm_bow_transform. What the device returns per feature is the triple DBoW2's
// transform(feature, id, w, &nid, levelsup) computes; the two std::maps are filled and normalised here, in feature order,
// by DBoW2's own BowVector / FeatureVector methods (:1147-1160, :1198) — every double is summed by the reference's code in
// the reference's order. tests/test_shim_bow_vs_reference_source.py runs these bodies against the reference's own two
// functions over DBoW2's own TemplatedVocabulary (CPU: ABI answered by the oracle; GPU: by liborbx.so).
#include "Frame.h"
#include "KeyFrame.h"
#include "ORBVocabulary.h"

#include <cstring>
#include <stdexcept>
#include <vector>

#include "orbm.h"
#include "orbx_thread_matcher.h"

namespace ORB_SLAM3 {

namespace {

// A derived class may name the protected members of TemplatedVocabulary (:405-427), and a pointer to a member named
// that way applies to any ORBVocabulary object.
struct VocabularyReader : public ORBVocabulary {
  struct Flat {
    std::vector<int32_t> child_offsets;
    std::vector<uint32_t> children, word_id;
    std::vector<uint8_t> descriptors;
    std::vector<double> weight;
  };
  static const std::vector<Node>& nodes(const ORBVocabulary& v) { return v.*(&VocabularyReader::m_nodes); }
  static Flat flatten(const ORBVocabulary& v) {
    const std::vector<Node>& nd = nodes(v);
    const size_t N = nd.size();
    Flat f;
    f.child_offsets.assign(N + 1, 0);
    f.word_id.assign(N, 0);
    f.weight.assign(N, 0.0);
    f.descriptors.assign(N * 32, 0);
    for (size_t i = 0; i < N; i++) {
      f.child_offsets[i + 1] = f.child_offsets[i] + (int32_t)nd[i].children.size();
      for (DBoW2::NodeId c : nd[i].children) f.children.push_back((uint32_t)c);
      f.word_id[i] = (uint32_t)nd[i].word_id;
      f.weight[i] = nd[i].weight;
      if (i > 0 && !nd[i].descriptor.empty()) memcpy(&f.descriptors[i * 32], nd[i].descriptor.ptr(0), 32);  // FORB: 1 x 32
    }
    return f;
  }
  // cheap identity of a tree: its size and a few sampled weights (one vocabulary lives as long as the System does, but
  // an address alone could be reused)
  static double fingerprint(const ORBVocabulary& v) {
    const std::vector<Node>& nd = nodes(v);
    double s = (double)nd.size();
    for (size_t k = 1; k <= 16 && !nd.empty(); k++) s += nd[(nd.size() - 1) * k / 16].weight * (double)k;
    return s;
  }
  static bool must_normalize(const ORBVocabulary& v, DBoW2::LNorm& norm) {
    return (v.*(&VocabularyReader::m_scoring_object))->mustNormalize(norm);
  }
};

// the calling thread's matcher with `voc` resident on its device
orbm_matcher* MatcherWithVocabulary(const ORBVocabulary& voc) {
  thread_local const ORBVocabulary* loaded = nullptr;
  thread_local orbm_matcher* loaded_on = nullptr;
  thread_local double loaded_print = 0;
  orbm_matcher* m = OrbxThreadMatcher();
  const double print = VocabularyReader::fingerprint(voc);
  if (loaded == &voc && loaded_on == m && loaded_print == print) return m;
  const VocabularyReader::Flat f = VocabularyReader::flatten(voc);
  orbx_vocabulary view;
  view.n_nodes = (int32_t)f.word_id.size();
  view.depth = voc.getDepthLevels();
  view.child_offsets = f.child_offsets.data();
  view.children = f.children.data();
  view.descriptors = f.descriptors.data();
  view.word_id = f.word_id.data();
  view.weight = f.weight.data();
  if (orbm_set_vocabulary(m, &view) != ORBX_OK) throw std::runtime_error(orbm_last_error(m));
  loaded = &voc;
  loaded_on = m;
  loaded_print = print;
  return m;
}

// TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) (:1126-1200) for the rows of D
void TransformOnDevice(const ORBVocabulary& voc, const cv::Mat& D, DBoW2::BowVector& v, DBoW2::FeatureVector& fv,
                       int levelsup) {
  v.clear();
  fv.clear();
  if (voc.empty()) return;
  const int n = D.rows;
  std::vector<uint8_t> rows((size_t)n * 32);
  for (int i = 0; i < n; i++) memcpy(&rows[(size_t)i * 32], D.ptr(i), 32);
  std::vector<uint32_t> word(n), node(n);
  std::vector<double> weight(n);
  if (n > 0) {
    orbm_matcher* m = MatcherWithVocabulary(voc);
    if (orbm_bow_transform(m, rows.data(), n, levelsup, word.data(), weight.data(), node.data()) != ORBX_OK)
      throw std::runtime_error(orbm_last_error(m));
  }
  const DBoW2::WeightingType wt = voc.getWeightingType();
  const bool sums = wt == DBoW2::TF || wt == DBoW2::TF_IDF;
  for (int i = 0; i < n; i++) {
    if (!(weight[i] > 0)) continue;  // a stopped word is filed nowhere (:1154, :1181)
    if (sums)
      v.addWeight(word[i], weight[i]);
    else
      v.addIfNotExist(word[i], weight[i]);
    fv.addFeature(node[i], (unsigned int)i);
  }
  DBoW2::LNorm norm;
  const bool must = VocabularyReader::must_normalize(voc, norm);
  if (sums && !v.empty() && !must) {
    const double nd = v.size();
    for (DBoW2::BowVector::iterator it = v.begin(); it != v.end(); ++it) it->second /= nd;
  }
  if (must) v.normalize(norm);
}

}  // namespace

void Frame::ComputeBoW() {
  if (mBowVec.empty()) TransformOnDevice(*mpORBvocabulary, mDescriptors, mBowVec, mFeatVec, 4);
}

void KeyFrame::ComputeBoW() {
  if (mBowVec.empty() || mFeatVec.empty()) TransformOnDevice(*mpORBvocabulary, mDescriptors, mBowVec, mFeatVec, 4);
}

}  // namespace ORB_SLAM3
