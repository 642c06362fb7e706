// shim/ORBmatcher_next_orbx.cc — the SURVEY.md §8(f) "next" rows that are built so far, forwarding to the orbm C ABI:
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)            src/ORBmatcher.cc:230-404
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)         :766-884
//   ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight)         :1108-1275
//   Frame::AssignFeaturesToGrid() (both rigs)                                 src/Frame.cc:520-547
//   MapPoint::ComputeDistinctiveDescriptors() for a batch of points           src/MapPoint.cc:372-441
//
// COMPILES ONLY INSIDE THE REFERENCE TREE (needs the reference headers and their OpenCV / Eigen / Sophus / DBoW2
// dependencies). Both SearchByBoW overloads and AssignFeaturesToGrid cover both rigs; the others are the pinhole
// forms; Fuse(KeyFrame*, vpMapPoints, th, bRight) also covers two-camera KeyFrames and bRight. As in ORBmatcher_orbx.cc the shim only flattens the pointer graph and scatters the answers back; every
// float that decides a match comes from the reference's own expressions on the host (projection) or from the device
// with the same non-fused FP32 operations.
#include "ORBmatcher.h"
#include "orbm.h"
#include "orbx_thread_matcher.h"

namespace ORB_SLAM3 {

namespace {

// DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) -> CSR, plus the per-feature flag the search reads
struct BowFlat {
  std::vector<uint8_t> has_mp;
  std::vector<uint32_t> ids, idx;
  std::vector<int32_t> off{0};
  orbx_keyframe_view v;
  template <class Keys>
  BowFlat(int N, const Keys& keys, const cv::Mat& desc, const DBoW2::FeatureVector& fv,
          const std::vector<float>& scale, const std::vector<float>& sigma2) : has_mp(N, 0) {
    for (const auto& node : fv) {
      ids.push_back(node.first);
      idx.insert(idx.end(), node.second.begin(), node.second.end());
      off.push_back((int32_t)idx.size());
    }
    v.n = N;
    v.kps = reinterpret_cast<const orbx_kp*>(keys.data());
    v.desc = desc.data;
    v.u_right = nullptr;
    v.has_mappoint = has_mp.data();
    v.featvec = orbx_featvec{(int32_t)ids.size(), ids.data(), off.data(), idx.data()};
    v.scale_factors = scale.data();
    v.level_sigma2 = sigma2.data();
    v.n_levels = (int32_t)scale.size();
  }
};
}  // namespace

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
  const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
  BowFlat kf(pKF->N, pKF->mvKeysUn, pKF->mDescriptors, pKF->mFeatVec, pKF->mvScaleFactors, pKF->mvLevelSigma2);
  for (int i = 0; i < pKF->N; i++) kf.has_mp[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();  // :262-264
  std::vector<int32_t> mf(F.N, -1);
  int32_t nmatches = 0;
  if (F.Nleft == -1) {
    BowFlat fr(F.N, F.mvKeys, F.mDescriptors, F.mFeatVec, F.mvScaleFactors, F.mvLevelSigma2);  // angles: F.mvKeys
    if (orbm_search_by_bow(OrbxThreadMatcher(), &kf.v, &fr.v, mfNNratio, mbCheckOrientation, mf.data(), &nmatches) != ORBX_OK)
      throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  } else {
    // two-camera rig (:274-365): rows >= Nleft are the right camera's; the keypoint of a row is mvKeys[i] or
    // mvKeysRight[i - Nleft] on both sides (:323-335, :352-362) — flattened here, left rows first
    auto both = [](const std::vector<cv::KeyPoint>& l, const std::vector<cv::KeyPoint>& r, int nl) {
      std::vector<cv::KeyPoint> k(l.begin(), l.begin() + nl);
      k.insert(k.end(), r.begin(), r.end());
      return k;
    };
    const std::vector<cv::KeyPoint> keysF = both(F.mvKeys, F.mvKeysRight, F.Nleft);
    const std::vector<cv::KeyPoint> keysK =
        pKF->mpCamera2 ? both(pKF->mvKeys, pKF->mvKeysRight, pKF->NLeft == -1 ? pKF->N : pKF->NLeft) : pKF->mvKeysUn;
    kf.v.kps = reinterpret_cast<const orbx_kp*>(keysK.data());
    BowFlat fr(F.N, keysF, F.mDescriptors, F.mFeatVec, F.mvScaleFactors, F.mvLevelSigma2);
    if (orbm_search_by_bow_fisheye(OrbxThreadMatcher(), &kf.v, &fr.v, F.Nleft, mfNNratio, mbCheckOrientation, mf.data(),
                                   &nmatches) != ORBX_OK)
      throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  }
  vpMapPointMatches.assign(F.N, static_cast<MapPoint*>(NULL));  // :235
  for (int i = 0; i < F.N; i++)
    if (mf[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[mf[i]];
  return nmatches;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
  const std::vector<MapPoint*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
  BowFlat k1(pKF1->N, pKF1->mvKeysUn, pKF1->mDescriptors, pKF1->mFeatVec, pKF1->mvScaleFactors, pKF1->mvLevelSigma2);
  BowFlat k2(pKF2->N, pKF2->mvKeysUn, pKF2->mDescriptors, pKF2->mFeatVec, pKF2->mvScaleFactors, pKF2->mvLevelSigma2);
  // two-camera KeyFrames: rows past mvKeysUn (the right camera's) are skipped on both sides (:799-801, :816-818) —
  // the same effect as "no MapPoint"; the angle arrays are padded so that every row has an entry
  const int un1 = pKF1->NLeft != -1 ? (int)pKF1->mvKeysUn.size() : pKF1->N;
  const int un2 = pKF2->NLeft != -1 ? (int)pKF2->mvKeysUn.size() : pKF2->N;
  std::vector<cv::KeyPoint> pad1, pad2;
  if (un1 < pKF1->N) {
    pad1 = pKF1->mvKeysUn;
    pad1.resize(pKF1->N);
    k1.v.kps = reinterpret_cast<const orbx_kp*>(pad1.data());
  }
  if (un2 < pKF2->N) {
    pad2 = pKF2->mvKeysUn;
    pad2.resize(pKF2->N);
    k2.v.kps = reinterpret_cast<const orbx_kp*>(pad2.data());
  }
  for (int i = 0; i < pKF1->N; i++) k1.has_mp[i] = i < un1 && mp1[i] && !mp1[i]->isBad();  // :802-804
  for (int i = 0; i < pKF2->N; i++) k2.has_mp[i] = i < un2 && mp2[i] && !mp2[i]->isBad();  // :821-825
  std::vector<int32_t> m12(pKF1->N, -1);
  int32_t nmatches = 0;
  if (orbm_search_by_bow_kf(OrbxThreadMatcher(), &k1.v, &k2.v, mfNNratio, mbCheckOrientation, m12.data(), &nmatches) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  vpMatches12.assign(mp1.size(), static_cast<MapPoint*>(NULL));  // :779-780
  for (int i = 0; i < pKF1->N; i++)
    if (m12[i] >= 0) vpMatches12[i] = mp2[m12[i]];
  return nmatches;
}

int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th, const bool bRight) {
  // Host part = :1116-1192 verbatim: camera, pose and centre of the searched camera (the right one with bRight), which
  // points take part, uv, ur, level. The projection stays with the reference's own camera object (mpCamera2 is a
  // KannalaBrandt8 model on two-camera rigs).
  GeometricCamera* pCamera = bRight ? pKF->mpCamera2 : pKF->mpCamera;
  const Sophus::SE3f Tcw = bRight ? pKF->GetRightPose() : pKF->GetPose();
  const Eigen::Vector3f Ow = bRight ? pKF->GetRightCameraCenter() : pKF->GetCameraCenter();
  const float bf = pKF->mbf;
  std::vector<int> src;
  std::vector<float> u, v, ur, radius;
  std::vector<int32_t> lmin, lmax;
  std::vector<uint8_t> desc;
  for (int i = 0; i < (int)vpMapPoints.size(); i++) {
    MapPoint* pMP = vpMapPoints[i];
    if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
    const Eigen::Vector3f p3Dw = pMP->GetWorldPos(), p3Dc = Tcw * p3Dw;
    if (p3Dc(2) < 0.0f) continue;
    const float invz = 1 / p3Dc(2);
    const Eigen::Vector2f uv = pCamera->project(p3Dc);
    if (!pKF->IsInImage(uv(0), uv(1))) continue;
    const Eigen::Vector3f PO = p3Dw - Ow;
    const float dist3D = PO.norm();
    if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
    if (PO.dot(pMP->GetNormal()) < 0.5 * dist3D) continue;
    const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
    src.push_back(i);
    u.push_back(uv(0));
    v.push_back(uv(1));
    ur.push_back(uv(0) - bf * invz);
    radius.push_back(th * pKF->mvScaleFactors[nPredictedLevel]);
    lmin.push_back(nPredictedLevel - 1);
    lmax.push_back(nPredictedLevel);
    const cv::Mat d = pMP->GetDescriptor();
    desc.insert(desc.end(), d.data, d.data + 32);
  }
  // KeyFrame view of the searched camera (:1200-1201, :1219-1221, :1247): pinhole rigs — mvKeysUn, mGrid, every row of
  // mDescriptors; two-camera rigs — mvKeys / mGrid / rows [0, NLeft), or with bRight mvKeysRight / mGridRight / rows
  // [NLeft, N), the fused row being the local index + NLeft. mvuRight is indexed with the LOCAL index (:1227).
  const bool two = pKF->NLeft != -1;
  const std::vector<cv::KeyPoint>& keys = !two ? pKF->mvKeysUn : (!bRight ? pKF->mvKeys : pKF->mvKeysRight);
  const int row0 = two && bRight ? pKF->NLeft : 0, nk = (int)keys.size();
  std::vector<int32_t> off(pKF->mnGridCols * pKF->mnGridRows + 1, 0), items;
  for (int c = 0; c < pKF->mnGridCols; c++)
    for (int r = 0; r < pKF->mnGridRows; r++) {
      // accessors to add next to mGrid / mGridRight (include/KeyFrame.h)
      const std::vector<size_t>& cell = two && bRight ? pKF->GetGridCellRight(c, r) : pKF->GetGridCell(c, r);
      off[c * pKF->mnGridRows + r + 1] = off[c * pKF->mnGridRows + r] + (int32_t)cell.size();
      for (size_t k : cell) items.push_back((int32_t)k);
    }
  std::vector<uint8_t> occupied(nk, 0);
  std::vector<float> ur_local(nk, -1.f);
  for (int k = 0; k < nk && k < (int)pKF->mvuRight.size(); k++) ur_local[k] = pKF->mvuRight[k];
  orbx_frame_view kv;
  kv.n = nk;
  kv.kps = reinterpret_cast<const orbx_kp*>(keys.data());
  kv.desc = pKF->mDescriptors.data + (size_t)row0 * 32;
  kv.u_right = ur_local.data();
  kv.occupied = occupied.data();
  kv.grid = orbx_grid{off.data(), items.data(), (float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv,
                      pKF->mfGridElementHeightInv};
  kv.scale_factors = pKF->mvScaleFactors.data();
  kv.n_levels = (int32_t)pKF->mvScaleFactors.size();
  std::vector<float> angle(src.size(), 0.f);
  std::vector<uint8_t> has_obs(src.size(), 0);
  orbx_projected pts{(int32_t)src.size(), u.data(), v.data(), ur.data(), radius.data(), lmin.data(), lmax.data(),
                     angle.data(), has_obs.data(), desc.data()};
  std::vector<int32_t> best_idx(src.size(), -1), best_dist(src.size(), 256);
  if (orbm_fuse_match(OrbxThreadMatcher(), &kv, pKF->mvInvLevelSigma2.data(), &pts, /*chi2_gate*/ 1, best_idx.data(),
                      best_dist.data()) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  int nFused = 0;  // :1261-1273, in point order
  for (size_t k = 0; k < src.size(); k++) {
    if (best_dist[k] > TH_LOW) continue;
    MapPoint* pMP = vpMapPoints[src[k]];
    if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;  // an earlier Replace may have moved it in (:1141-1147)
    const int bestIdx = best_idx[k] + row0;  // :1247
    MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) {
        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
        else pMPinKF->Replace(pMP);
      }
    } else {
      pMP->AddObservation(pKF, bestIdx);
      pKF->AddMapPoint(pMP, bestIdx);
    }
    nFused++;
  }
  return nFused;
}

// Frame::AssignFeaturesToGrid() (src/Frame.cc:520-547): mGrid from mvKeysUn (Nleft == -1), or — two-camera rigs — mGrid
// from mvKeys and mGridRight from mvKeysRight with indices local to the right camera (:534-546)
void AssignFeaturesToGrid_orbx(Frame& F) {
  auto build = [](const std::vector<cv::KeyPoint>& keys, int n, std::vector<std::size_t> (&grid)[FRAME_GRID_COLS][FRAME_GRID_ROWS]) {
    std::vector<int32_t> off(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1), items(n > 0 ? n : 1);
    if (orbm_assign_features_to_grid(OrbxThreadMatcher(), reinterpret_cast<const orbx_kp*>(keys.data()), n, Frame::mnMinX,
                                     Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv,
                                     off.data(), items.data()) != ORBX_OK)
      throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
    for (int c = 0; c < FRAME_GRID_COLS; c++)
      for (int r = 0; r < FRAME_GRID_ROWS; r++)
        grid[c][r].assign(items.begin() + off[c * FRAME_GRID_ROWS + r], items.begin() + off[c * FRAME_GRID_ROWS + r + 1]);
  };
  if (F.Nleft == -1) {
    build(F.mvKeysUn, F.N, F.mGrid);
  } else {
    build(F.mvKeys, F.Nleft, F.mGrid);
    build(F.mvKeysRight, F.N - F.Nleft, F.mGridRight);
  }
}

// MapPoint::ComputeDistinctiveDescriptors() (src/MapPoint.cc:372-441) for every point LocalMapping touched in one go:
// lists[p] = the observed descriptors of point p as gathered by :386-405; returns the winning position per point.
std::vector<int32_t> DistinctiveDescriptors_orbx(const std::vector<std::vector<cv::Mat>>& lists) {
  std::vector<uint8_t> all;
  std::vector<int32_t> off{0};
  for (const auto& l : lists) {
    for (const cv::Mat& d : l) all.insert(all.end(), d.data, d.data + 32);
    off.push_back((int32_t)(all.size() / 32));
  }
  std::vector<int32_t> best(lists.size(), -1);
  if (orbm_distinctive_descriptors(OrbxThreadMatcher(), all.data(), off.data(), (int)lists.size(), best.data()) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  return best;  // mDescriptor = lists[p][best[p]].clone()                                                 :437-440
}

}  // namespace ORB_SLAM3
