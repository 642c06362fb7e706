// shim/ORBmatcher_sim3_orbx.cc — the remaining SURVEY.md §8(f) rank-3 bodies, forwarding to the orbm C ABI:
//   ORBmatcher::SearchByProjection(KeyFrame*, Sophus::Sim3f&, vpPoints, vpMatched, th, ratioHamming)       src/ORBmatcher.cc:406-506
//   ORBmatcher::SearchByProjection(KeyFrame*, Sophus::Sim3<float>&, vpPoints, vpPointsKFs, vpMatched, vpMatchedKF, ...)  :508-616
//   ORBmatcher::SearchBySim3(KeyFrame*, KeyFrame*, vpMatches12, S12, th)                                    :1392-1592
//   ORBmatcher::Fuse(KeyFrame*, Sophus::Sim3f&, vpPoints, th, vpReplacePoint)                               :1283-1390
//   ORBmatcher::SearchForInitialization(Frame&, Frame&, vbPrevMatched, vnMatches12, windowSize)            :618-764
//   ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)                  :1808-1918
//   MapPoint::ComputeDistinctiveDescriptors()                                                              src/MapPoint.cc:372-441
//
// COMPILES ONLY INSIDE THE REFERENCE TREE (or the stand-in world of oracle/ref_stubs, where the tests run these bodies
// against the reference's own methods). Pinhole rigs (NLeft == -1). As in the other shim files the host part — which
// points take part, the Sim3 / SE3 products, the camera projection, PredictScale — is the reference's own expressions
// evaluated by Eigen / Sophus on the host; the device does the window walk, the level test and the Hamming distances.
// No new entry point is needed: the Sim3 projections are orbm_search_by_projection_frame with every written keypoint
// closing for later points, SearchBySim3 is two gate-free orbm_fuse_match passes plus the agreement loop.
#include <climits>

#include "ORBmatcher.h"
#include "orbm.h"
#include "orbx_thread_matcher.h"

namespace ORB_SLAM3 {

namespace {
// KeyFrame::mGrid (std::vector<std::vector<std::vector<size_t>>>, include/KeyFrame.h) -> the CSR of orbx_grid
struct KeyFrameFlat {
  std::vector<int32_t> off, items;
  std::vector<uint8_t> occupied;
  orbx_frame_view v;
  explicit KeyFrameFlat(KeyFrame* pKF) : off(pKF->mnGridCols * pKF->mnGridRows + 1, 0), occupied(pKF->N, 0) {
    for (int c = 0; c < pKF->mnGridCols; c++)
      for (int r = 0; r < pKF->mnGridRows; r++) {
        const std::vector<size_t>& cell = pKF->GetGridCell(c, r);  // accessor to add next to mGrid (include/KeyFrame.h)
        off[c * pKF->mnGridRows + r + 1] = off[c * pKF->mnGridRows + r] + (int32_t)cell.size();
        for (size_t k : cell) items.push_back((int32_t)k);
      }
    v.n = pKF->N;
    v.kps = reinterpret_cast<const orbx_kp*>(pKF->mvKeysUn.data());
    v.desc = pKF->mDescriptors.data;
    v.u_right = nullptr;
    v.occupied = occupied.data();
    v.grid = orbx_grid{off.data(), items.data(), (float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv,
                       pKF->mfGridElementHeightInv};
    v.scale_factors = pKF->mvScaleFactors.data();
    v.n_levels = (int32_t)pKF->mvScaleFactors.size();
  }
};

// Frame::mGrid -> CSR (as in ORBmatcher_orbx.cc), with the occupancy rule left to the caller
struct FrameGridFlat {
  std::vector<int32_t> off, items;
  std::vector<uint8_t> occupied;
  orbx_frame_view v;
  explicit FrameGridFlat(const Frame& F) : off(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0), occupied(F.N, 0) {
    for (int c = 0; c < FRAME_GRID_COLS; c++)
      for (int r = 0; r < FRAME_GRID_ROWS; r++) {
        off[c * FRAME_GRID_ROWS + r + 1] = off[c * FRAME_GRID_ROWS + r] + (int32_t)F.mGrid[c][r].size();
        for (size_t k : F.mGrid[c][r]) items.push_back((int32_t)k);
      }
    v.n = F.N;
    v.kps = reinterpret_cast<const orbx_kp*>(F.mvKeysUn.data());
    v.desc = F.mDescriptors.data;
    v.u_right = nullptr;
    v.occupied = occupied.data();
    v.grid = orbx_grid{off.data(), items.data(), Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv,
                       Frame::mfGridElementHeightInv};
    v.scale_factors = F.mvScaleFactors.data();
    v.n_levels = (int32_t)F.mvScaleFactors.size();
  }
};

// the points of a projected search as they are collected on the host
struct Projected {
  std::vector<int> src;
  std::vector<float> u, v, radius, angle;
  std::vector<int32_t> lo, hi;
  std::vector<uint8_t> has_obs, desc;
  void add(int i, float uu, float vv, float r, int l0, int l1, float a, MapPoint* pMP) {
    src.push_back(i);
    u.push_back(uu);
    v.push_back(vv);
    radius.push_back(r);
    lo.push_back(l0);
    hi.push_back(l1);
    angle.push_back(a);
    has_obs.push_back(1);  // every point that is written closes its keypoint for the later ones
    const cv::Mat d = pMP->GetDescriptor();
    desc.insert(desc.end(), d.data, d.data + 32);
  }
  orbx_projected view() const {
    return orbx_projected{(int32_t)src.size(), u.data(), v.data(), nullptr, radius.data(), lo.data(), hi.data(),
                          angle.data(), has_obs.data(), desc.data()};
  }
};

// shared body of the two Sim3 SearchByProjection overloads: kfs / matched_kf are NULL for the first
int Sim3Projection(KeyFrame* pKF, const Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints,
                   const std::vector<KeyFrame*>* kfs, std::vector<MapPoint*>& vpMatched,
                   std::vector<KeyFrame*>* matched_kf, int th, float ratioHamming, bool own_projection) {
  const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
  Sophus::SE3f Tcw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
  Eigen::Vector3f Ow = Tcw.inverse().translation();
  std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
  spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
  Projected P;
  for (int iMP = 0, iendMP = vpPoints.size(); iMP < iendMP; iMP++) {  // :430-472 / :533-583 on the host
    MapPoint* pMP = vpPoints[iMP];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    Eigen::Vector3f p3Dw = pMP->GetWorldPos();
    Eigen::Vector3f p3Dc = Tcw * p3Dw;
    if (p3Dc(2) < 0.0) continue;
    float u, v;
    if (own_projection) {  // the second overload projects by hand (:548-554)
      const float invz = 1 / p3Dc(2);
      const float x = p3Dc(0) * invz;
      const float y = p3Dc(1) * invz;
      u = fx * x + cx;
      v = fy * y + cy;
    } else {
      const Eigen::Vector2f uv = pKF->mpCamera->project(p3Dc);
      u = uv(0);
      v = uv(1);
    }
    if (!pKF->IsInImage(u, v)) continue;
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    Eigen::Vector3f PO = p3Dw - Ow;
    const float dist = PO.norm();
    if (dist < minDistance || dist > maxDistance) continue;
    Eigen::Vector3f Pn = pMP->GetNormal();
    if (PO.dot(Pn) < 0.5 * dist) continue;
    int nPredictedLevel = pMP->PredictScale(dist, pKF);
    P.add(iMP, u, v, th * pKF->mvScaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, 0.f, pMP);
  }
  KeyFrameFlat kf(pKF);
  for (int i = 0; i < pKF->N; i++) kf.occupied[i] = vpMatched[i] != nullptr;  // if (vpMatched[idx]) continue;  :478
  const orbx_projected pts = P.view();
  std::vector<int32_t> assign(pKF->N, -1);
  int32_t nmatches = 0;
  // bestDist <= TH_LOW * ratioHamming (:499) for an integer bestDist
  const int max_dist = (int)std::floor(ORBmatcher::TH_LOW * ratioHamming);
  if (orbm_search_by_projection_frame(OrbxThreadMatcher(), &kf.v, &pts, max_dist, /*check_orientation*/ 0, assign.data(),
                                      &nmatches) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  for (int i = 0; i < pKF->N; i++)
    if (assign[i] >= 0) {
      vpMatched[i] = vpPoints[P.src[assign[i]]];                       // :500 / :609
      if (matched_kf) (*matched_kf)[i] = (*kfs)[P.src[assign[i]]];     // :610
    }
  return nmatches;
}
}  // namespace

int ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints,
                                   std::vector<MapPoint*>& vpMatched, int th, float ratioHamming) {
  return Sim3Projection(pKF, Scw, vpPoints, nullptr, vpMatched, nullptr, th, ratioHamming, false);
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3<float>& Scw, const std::vector<MapPoint*>& vpPoints,
                                   const std::vector<KeyFrame*>& vpPointsKFs, std::vector<MapPoint*>& vpMatched,
                                   std::vector<KeyFrame*>& vpMatchedKF, int th, float ratioHamming) {
  return Sim3Projection(pKF, Scw, vpPoints, &vpPointsKFs, vpMatched, &vpMatchedKF, th, ratioHamming, true);
}

int ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints, float th,
                     std::vector<MapPoint*>& vpReplacePoint) {
  Sophus::SE3f Tcw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
  Eigen::Vector3f Ow = Tcw.inverse().translation();
  const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
  Projected P;
  const int nPoints = vpPoints.size();
  for (int iMP = 0; iMP < nPoints; iMP++) {  // :1307-1342 on the host
    MapPoint* pMP = vpPoints[iMP];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    Eigen::Vector3f p3Dw = pMP->GetWorldPos();
    Eigen::Vector3f p3Dc = Tcw * p3Dw;
    if (p3Dc(2) < 0.0f) continue;
    const Eigen::Vector2f uv = pKF->mpCamera->project(p3Dc);
    if (!pKF->IsInImage(uv(0), uv(1))) continue;
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    Eigen::Vector3f PO = p3Dw - Ow;
    const float dist3D = PO.norm();
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    Eigen::Vector3f Pn = pMP->GetNormal();
    if (PO.dot(Pn) < 0.5 * dist3D) continue;
    const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
    P.add(iMP, uv(0), uv(1), th * pKF->mvScaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, 0.f, pMP);
  }
  KeyFrameFlat kf(pKF);
  const orbx_projected pts = P.view();
  std::vector<int32_t> best_idx(P.src.size(), -1), best_dist(P.src.size(), 256);
  if (orbm_fuse_match(OrbxThreadMatcher(), &kf.v, pKF->mvInvLevelSigma2.data(), &pts, /*chi2_gate*/ 0, best_idx.data(),
                      best_dist.data()) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  int nFused = 0;  // :1374-1385, in point order (AddMapPoint changes what a later point finds on the keypoint)
  for (size_t k = 0; k < P.src.size(); k++) {
    if (best_dist[k] > TH_LOW) continue;
    MapPoint* pMP = vpPoints[P.src[k]];
    MapPoint* pMPinKF = pKF->GetMapPoint(best_idx[k]);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) vpReplacePoint[P.src[k]] = pMPinKF;
    } else {
      pMP->AddObservation(pKF, best_idx[k]);
      pKF->AddMapPoint(pMP, best_idx[k]);
    }
    nFused++;
  }
  return nFused;
}

int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12,
                             const Sophus::Sim3f& S12, const float th) {
  const float &fx = pKF1->fx, &fy = pKF1->fy, &cx = pKF1->cx, &cy = pKF1->cy;
  Sophus::SE3f T1w = pKF1->GetPose();
  Sophus::SE3f T2w = pKF2->GetPose();
  Sophus::Sim3f S21 = S12.inverse();
  const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
  const int N1 = vpMapPoints1.size();
  const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
  const int N2 = vpMapPoints2.size();
  std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
  for (int i = 0; i < N1; i++) {  // :1418-1426
    MapPoint* pMP = vpMatches12[i];
    if (pMP) {
      vbAlreadyMatched1[i] = true;
      int idx2 = std::get<0>(pMP->GetIndexInKeyFrame(pKF2));
      if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
    }
  }
  // one direction: the MapPoints of `from` projected into `to` (:1432-1500 and :1503-1571 are the same loop mirrored)
  auto direction = [&](const std::vector<MapPoint*>& pts_from, const std::vector<bool>& done, const Sophus::SE3f& Tfw,
                       const Sophus::Sim3f& S_to_from, KeyFrame* to, std::vector<int>& match) {
    Projected P;
    for (int i = 0; i < (int)pts_from.size(); i++) {
      MapPoint* pMP = pts_from[i];
      if (!pMP || done[i]) continue;
      if (pMP->isBad()) continue;
      Eigen::Vector3f p3Dw = pMP->GetWorldPos();
      Eigen::Vector3f p3Dcf = Tfw * p3Dw;
      Eigen::Vector3f p3Dct = S_to_from * p3Dcf;
      if (p3Dct(2) < 0.0) continue;
      const float invz = 1.0 / p3Dct(2);
      const float x = p3Dct(0) * invz;
      const float y = p3Dct(1) * invz;
      const float u = fx * x + cx;
      const float v = fy * y + cy;
      if (!to->IsInImage(u, v)) continue;
      const float maxDistance = pMP->GetMaxDistanceInvariance();
      const float minDistance = pMP->GetMinDistanceInvariance();
      const float dist3D = p3Dct.norm();
      if (dist3D < minDistance || dist3D > maxDistance) continue;
      const int nPredictedLevel = pMP->PredictScale(dist3D, to);
      P.add(i, u, v, th * to->mvScaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, 0.f, pMP);
    }
    KeyFrameFlat kf(to);
    const orbx_projected pts = P.view();
    std::vector<int32_t> best_idx(P.src.size(), -1), best_dist(P.src.size(), 256);
    if (orbm_fuse_match(OrbxThreadMatcher(), &kf.v, to->mvInvLevelSigma2.data(), &pts, /*chi2_gate*/ 0, best_idx.data(),
                        best_dist.data()) != ORBX_OK)
      throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
    for (size_t k = 0; k < P.src.size(); k++)
      if (best_idx[k] >= 0 && best_dist[k] <= TH_HIGH) match[P.src[k]] = best_idx[k];  // :1497-1499
  };
  std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
  direction(vpMapPoints1, vbAlreadyMatched1, T1w, S21, pKF2, vnMatch1);
  direction(vpMapPoints2, vbAlreadyMatched2, T2w, S12, pKF1, vnMatch2);
  int nFound = 0;  // :1574-1588
  for (int i1 = 0; i1 < N1; i1++) {
    int idx2 = vnMatch1[i1];
    if (idx2 >= 0) {
      int idx1 = vnMatch2[idx2];
      if (idx1 == i1) {
        vpMatches12[i1] = vpMapPoints2[idx2];
        nFound++;
      }
    }
  }
  return nFound;
}

int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                        std::vector<int>& vnMatches12, int windowSize) {
  vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);
  const int n1 = (int)F1.mvKeysUn.size();
  if (n1 == 0) return 0;
  FrameGridFlat f1(F1), f2(F2);  // F1's grid is not read by the call; F2's is
  static_assert(sizeof(cv::Point2f) == 8, "vbPrevMatched crosses the ABI as float pairs");
  std::vector<int32_t> m12(n1, -1);
  int32_t nmatches = 0;
  if (orbm_search_for_initialization(OrbxThreadMatcher(), &f1.v, &f2.v, reinterpret_cast<const float*>(vbPrevMatched.data()),
                                     windowSize, mfNNratio, mbCheckOrientation, m12.data(), &nmatches) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  for (int i = 0; i < n1; i++) vnMatches12[i] = m12[i];
  for (size_t i1 = 0, iend1 = vnMatches12.size(); i1 < iend1; i1++)  // :758-761
    if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;
  return nmatches;
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound,
                                   const float th, const int ORBdist) {
  const Sophus::SE3f Tcw = CurrentFrame.GetPose();
  Eigen::Vector3f Ow = Tcw.inverse().translation();
  const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
  Projected P;
  for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {  // :1826-1853 on the host
    MapPoint* pMP = vpMPs[i];
    if (!pMP) continue;
    if (pMP->isBad() || sAlreadyFound.count(pMP)) continue;
    Eigen::Vector3f x3Dw = pMP->GetWorldPos();
    Eigen::Vector3f x3Dc = Tcw * x3Dw;
    const Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
    if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
    if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
    Eigen::Vector3f PO = x3Dw - Ow;
    float dist3D = PO.norm();
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
    P.add((int)i, uv(0), uv(1), th * CurrentFrame.mvScaleFactors[nPredictedLevel], nPredictedLevel - 1,
          nPredictedLevel + 1, pKF->mvKeysUn[i].angle, pMP);
  }
  FrameGridFlat f(CurrentFrame);
  for (int i = 0; i < CurrentFrame.N; i++) f.occupied[i] = CurrentFrame.mvpMapPoints[i] != nullptr;  // ANY point blocks, :1862
  const orbx_projected pts = P.view();
  std::vector<int32_t> assign(CurrentFrame.N, -1);
  int32_t nmatches = 0;
  if (orbm_search_by_projection_frame(OrbxThreadMatcher(), &f.v, &pts, ORBdist, mbCheckOrientation, assign.data(),
                                      &nmatches) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  for (int i = 0; i < CurrentFrame.N; i++)
    if (assign[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[P.src[assign[i]]];  // :1875 (rotation rejects come back as -1)
  return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors() (src/MapPoint.cc:372-441): the gathering of the observed descriptors stays,
// the N x N distances and the median selection go to the device. LocalMapping calls it per point; a caller that has
// several points at hand uses DistinctiveDescriptors_orbx (ORBmatcher_next_orbx.cc) for all of them in one call.
void MapPoint::ComputeDistinctiveDescriptors() {
  std::vector<cv::Mat> vDescriptors;
  std::map<KeyFrame*, std::tuple<int, int>> observations;
  {
    std::unique_lock<std::mutex> lock1(mMutexFeatures);
    if (mbBad) return;
    observations = mObservations;
  }
  if (observations.empty()) return;
  vDescriptors.reserve(observations.size());
  for (auto mit = observations.begin(), mend = observations.end(); mit != mend; mit++) {
    KeyFrame* pKF = mit->first;
    if (!pKF->isBad()) {
      std::tuple<int, int> indexes = mit->second;
      int leftIndex = std::get<0>(indexes), rightIndex = std::get<1>(indexes);
      if (leftIndex != -1) vDescriptors.push_back(pKF->mDescriptors.row(leftIndex));
      if (rightIndex != -1) vDescriptors.push_back(pKF->mDescriptors.row(rightIndex));
    }
  }
  if (vDescriptors.empty()) return;
  std::vector<uint8_t> all;
  for (const cv::Mat& d : vDescriptors) all.insert(all.end(), d.data, d.data + 32);
  const int32_t off[2] = {0, (int32_t)vDescriptors.size()};
  int32_t best = -1;
  if (orbm_distinctive_descriptors(OrbxThreadMatcher(), all.data(), off, 1, &best) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    mDescriptor = vDescriptors[best].clone();  // :437-440
  }
}

}  // namespace ORB_SLAM3
