// shim/ORBmatcher_orbx.cc — bodies of the three ORBmatcher searches on the hot path, forwarding to the orbm C ABI.
//
// COMPILES ONLY INSIDE THE REFERENCE TREE (needs include/ORBmatcher.h, Frame.h, KeyFrame.h, MapPoint.h and their
// OpenCV / Eigen / Sophus / DBoW2 dependencies). Replace the bodies of
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)   src/ORBmatcher.cc:42-221
//   ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)                                 :1594-1806
//   ORBmatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, vMatchedPairs, bOnlyStereo, bCoarse)   :886-1106
// with the ones below. The local-map SearchByProjection and SearchForTriangulation cover both rigs (the two-camera
// forms: right-camera twin :148-217; camera-pair selection :1007-1043 with the epipolar test left to the reference's
// KannalaBrandt8 object); the frame-to-frame SearchByProjection covers both rigs too (two-camera frames: one
// orbm_search_by_projection_frame_decisions call per camera + the shared rotation histogram here).
// The shim's only job is flattening the pointer graph into the SoA / CSR views of include/orbx_types.h and scattering
// the answers back; every float that decides a match is computed by the reference's own expressions on the host
// (projection, radius) or by the device with the same non-fused FP32 operations.
#include "ORBmatcher.h"
#include "orbm.h"
#include "orbx_thread_matcher.h"

namespace ORB_SLAM3 {

namespace {

// Frame::mGrid[64][48] (std::vector<size_t> per cell, include/Frame.h:279) -> CSR; cell id = col * 48 + row
struct GridCSR {
  std::vector<int32_t> offsets, items;
  explicit GridCSR(const Frame& F) : offsets(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0) {
    for (int c = 0; c < FRAME_GRID_COLS; c++)
      for (int r = 0; r < FRAME_GRID_ROWS; r++) {
        offsets[c * FRAME_GRID_ROWS + r + 1] = offsets[c * FRAME_GRID_ROWS + r] + (int32_t)F.mGrid[c][r].size();
        for (size_t i : F.mGrid[c][r]) items.push_back((int32_t)i);
      }
  }
};

struct FrameFlat {
  GridCSR grid;
  std::vector<uint8_t> occupied;
  orbx_frame_view v;
  explicit FrameFlat(const Frame& F) : grid(F), occupied(F.N) {
    for (int i = 0; i < F.N; i++)  // the skip rule of :92-93
      occupied[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;
    v.n = F.N;
    v.kps = reinterpret_cast<const orbx_kp*>(F.mvKeysUn.data());
    v.desc = F.mDescriptors.data;
    v.u_right = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
    v.occupied = occupied.data();
    v.grid = orbx_grid{grid.offsets.data(), grid.items.data(), Frame::mnMinX, Frame::mnMinY,
                       Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    v.scale_factors = F.mvScaleFactors.data();
    v.n_levels = (int32_t)F.mvScaleFactors.size();
  }
};
}  // namespace

namespace {
// the two-camera form (F.Nleft != -1, KannalaBrandt8 rigs): both searches per point and the stereo partners,
// src/ORBmatcher.cc:42-221 incl. :148-217 -> orbm_search_by_projection_map_fisheye
int SearchByProjectionTwoCameras(Frame& F, const std::vector<MapPoint*>& vpMapPoints, float th, float nnratio,
                                 bool bFarPoints, float thFarPoints) {
  const int M = (int)vpMapPoints.size(), NL = F.Nleft, NR = F.N - F.Nleft;
  std::vector<uint8_t> inL(M), inR(M), has_obs(M), desc((size_t)M * 32);
  std::vector<float> px(M), py(M), pxr(M), pyr(M), vcos(M), vcosr(M), depth(M), unused(M, 0.f);
  std::vector<int32_t> level(M), levelr(M);
  for (int i = 0; i < M; i++) {  // MapPoint tracking scratch of both cameras, include/MapPoint.h:172-180
    MapPoint* p = vpMapPoints[i];
    const bool good = !p->isBad();
    inL[i] = p->mbTrackInView && good;
    inR[i] = p->mbTrackInViewR && good;
    px[i] = p->mTrackProjX; py[i] = p->mTrackProjY; pxr[i] = p->mTrackProjXR; pyr[i] = p->mTrackProjYR;
    level[i] = p->mnTrackScaleLevel; levelr[i] = p->mnTrackScaleLevelR;
    vcos[i] = p->mTrackViewCos; vcosr[i] = p->mTrackViewCosR; depth[i] = p->mTrackDepth;
    has_obs[i] = p->Observations() > 0;
    if (inL[i] || inR[i]) memcpy(&desc[(size_t)i * 32], p->GetDescriptor().data, 32);
  }
  auto csr = [](const std::vector<std::size_t> (&grid)[FRAME_GRID_COLS][FRAME_GRID_ROWS], std::vector<int32_t>& off,
                std::vector<int32_t>& items) {
    off.assign(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0);
    for (int c = 0; c < FRAME_GRID_COLS; c++)
      for (int r = 0; r < FRAME_GRID_ROWS; r++) {
        off[c * FRAME_GRID_ROWS + r + 1] = off[c * FRAME_GRID_ROWS + r] + (int32_t)grid[c][r].size();
        for (std::size_t k : grid[c][r]) items.push_back((int32_t)k);
      }
  };
  std::vector<int32_t> offL, itemsL, offR, itemsR;
  csr(F.mGrid, offL, itemsL);
  csr(F.mGridRight, offR, itemsR);
  std::vector<uint8_t> occupied(F.N);
  for (int i = 0; i < F.N; i++) occupied[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;  // :92-93, :183-185
  static_assert(sizeof(int) == sizeof(int32_t), "mvLeftToRightMatch crosses the ABI as int32");
  orbx_fisheye_view fv;
  fv.n_left = NL;
  fv.n_right = NR;
  fv.kps_left = reinterpret_cast<const orbx_kp*>(F.mvKeys.data());
  fv.kps_right = reinterpret_cast<const orbx_kp*>(F.mvKeysRight.data());
  fv.desc = F.mDescriptors.data;
  fv.occupied = occupied.data();
  fv.grid_left = orbx_grid{offL.data(), itemsL.data(), Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv,
                           Frame::mfGridElementHeightInv};
  fv.grid_right = orbx_grid{offR.data(), itemsR.data(), Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv,
                            Frame::mfGridElementHeightInv};
  fv.left_to_right = F.mvLeftToRightMatch.data();
  fv.right_to_left = F.mvRightToLeftMatch.data();
  fv.scale_factors = F.mvScaleFactors.data();
  fv.n_levels = (int32_t)F.mvScaleFactors.size();
  orbx_mappoints mp{M, inL.data(), px.data(), py.data(), unused.data(), level.data(), vcos.data(), depth.data(),
                    has_obs.data(), desc.data()};
  orbx_mappoints_right mr{inR.data(), pxr.data(), pyr.data(), levelr.data(), vcosr.data()};
  std::vector<int32_t> assign(F.N, -1);
  int32_t nmatches = 0;
  if (orbm_search_by_projection_map_fisheye(OrbxThreadMatcher(), &fv, &mp, &mr, th, nnratio, bFarPoints, thFarPoints,
                                            assign.data(), &nmatches) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  for (int i = 0; i < F.N; i++)
    if (assign[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[assign[i]];  // :130-137, :203-211
  return nmatches;
}
}  // namespace

int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th,
                                   const bool bFarPoints, const float thFarPoints) {
  if (F.Nleft != -1) return SearchByProjectionTwoCameras(F, vpMapPoints, th, mfNNratio, bFarPoints, thFarPoints);
  const int M = (int)vpMapPoints.size();
  std::vector<uint8_t> in_view(M), has_obs(M), desc((size_t)M * 32);
  std::vector<float> px(M), py(M), pxr(M), vcos(M), depth(M);
  std::vector<int32_t> level(M);
  for (int i = 0; i < M; i++) {  // MapPoint tracking scratch, include/MapPoint.h:172-180
    MapPoint* p = vpMapPoints[i];
    in_view[i] = p->mbTrackInView && !p->isBad();
    px[i] = p->mTrackProjX; py[i] = p->mTrackProjY; pxr[i] = p->mTrackProjXR;
    level[i] = p->mnTrackScaleLevel; vcos[i] = p->mTrackViewCos; depth[i] = p->mTrackDepth;
    has_obs[i] = p->Observations() > 0;
    if (in_view[i]) memcpy(&desc[(size_t)i * 32], p->GetDescriptor().data, 32);
  }
  FrameFlat ff(F);
  orbx_mappoints mp{M, in_view.data(), px.data(), py.data(), pxr.data(), level.data(), vcos.data(), depth.data(),
                    has_obs.data(), desc.data()};
  std::vector<int32_t> assign(F.N, -1);
  int32_t nmatches = 0;
  if (orbm_search_by_projection_map(OrbxThreadMatcher(), &ff.v, &mp, th, mfNNratio, bFarPoints, thFarPoints, assign.data(),
                                    &nmatches) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  for (int i = 0; i < F.N; i++)
    if (assign[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[assign[i]];  // :130
  return nmatches;
}

namespace {
// One camera of a two-camera frame as an orbx_frame_view: its keypoints, its grid, its rows of mDescriptors and
// mvpMapPoints (the skip rule of :1663-1665 / :1736-1740), no stereo gate (:1667 needs Nleft == -1).
struct CameraFlat {
  std::vector<int32_t> off, items;
  std::vector<uint8_t> occupied;
  orbx_frame_view v;
  CameraFlat(const Frame& F, bool right) : off(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0) {
    const std::vector<cv::KeyPoint>& keys = right ? F.mvKeysRight : F.mvKeys;
    const int row0 = right ? F.Nleft : 0, n = right ? F.N - F.Nleft : F.Nleft;
    for (int c = 0; c < FRAME_GRID_COLS; c++)
      for (int r = 0; r < FRAME_GRID_ROWS; r++) {
        const std::vector<std::size_t>& cell = right ? F.mGridRight[c][r] : F.mGrid[c][r];
        off[c * FRAME_GRID_ROWS + r + 1] = off[c * FRAME_GRID_ROWS + r] + (int32_t)cell.size();
        for (std::size_t i : cell) items.push_back((int32_t)i);
      }
    occupied.resize(n);
    for (int i = 0; i < n; i++)
      occupied[i] = F.mvpMapPoints[row0 + i] && F.mvpMapPoints[row0 + i]->Observations() > 0;
    v.n = n;
    v.kps = reinterpret_cast<const orbx_kp*>(keys.data());
    v.desc = F.mDescriptors.data + (size_t)row0 * 32;
    v.u_right = nullptr;
    v.occupied = occupied.data();
    v.grid = orbx_grid{off.data(), items.data(), Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv,
                       Frame::mfGridElementHeightInv};
    v.scale_factors = F.mvScaleFactors.data();
    v.n_levels = (int32_t)F.mvScaleFactors.size();
  }
};
}  // namespace

// The two-camera form (CurrentFrame.Nleft != -1, :1594-1806 with the right-camera block :1708-1780). The two cameras
// write disjoint rows of mvpMapPoints, so each is one orbm_search_by_projection_frame_decisions call (the serial order
// dependence of :1663-1665 is resolved on the device per camera); what they share — nmatches and ONE rotation
// histogram — is put together here in the reference's point order with the reference's own expressions.
namespace {
int SearchByProjectionTwoCameraFrames(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono,
                                      const bool mbCheckOrientation, std::vector<int>* rotHist /* [HISTO_LENGTH] */) {
  const int TH_HIGH = ORBmatcher::TH_HIGH, HISTO_LENGTH = ORBmatcher::HISTO_LENGTH;
  const Sophus::SE3f Tcw = CurrentFrame.GetPose();
  const Eigen::Vector3f twc = Tcw.inverse().translation();
  const Sophus::SE3f Tlw = LastFrame.GetPose();
  const Eigen::Vector3f tlc = Tlw * twc;
  const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;
  const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;
  const Sophus::SE3f Trl = CurrentFrame.GetRelativePoseTrl();
  std::vector<float> u, v, ur2, vr2, radius, angle;
  std::vector<int32_t> lo, hi, src;
  std::vector<uint8_t> has_obs, desc;
  for (int i = 0; i < LastFrame.N; i++) {
    MapPoint* pMP = LastFrame.mvpMapPoints[i];
    if (!pMP || LastFrame.mvbOutlier[i]) continue;
    const Eigen::Vector3f x3Dc = Tcw * pMP->GetWorldPos();
    const float invzc = 1.0 / x3Dc(2);
    if (invzc < 0) continue;
    const Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
    if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
    if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
    const int nLastOctave = (LastFrame.Nleft == -1 || i < LastFrame.Nleft)
                                ? LastFrame.mvKeys[i].octave
                                : LastFrame.mvKeysRight[i - LastFrame.Nleft].octave;
    const Eigen::Vector3f x3Dr = Trl * x3Dc;                                   // :1709-1710 (mpCamera, as the reference)
    const Eigen::Vector2f uvr = CurrentFrame.mpCamera->project(x3Dr);
    u.push_back(uv(0)); v.push_back(uv(1));
    ur2.push_back(uvr(0)); vr2.push_back(uvr(1));
    radius.push_back(th * CurrentFrame.mvScaleFactors[nLastOctave]);
    lo.push_back(bForward ? nLastOctave : (bBackward ? 0 : nLastOctave - 1));
    hi.push_back(bForward ? -1 : (bBackward ? nLastOctave : nLastOctave + 1));
    const cv::KeyPoint& kpLF = (LastFrame.Nleft == -1) ? LastFrame.mvKeysUn[i]
                               : (i < LastFrame.Nleft) ? LastFrame.mvKeys[i]
                                                       : LastFrame.mvKeysRight[i - LastFrame.Nleft];
    angle.push_back(kpLF.angle);
    has_obs.push_back(pMP->Observations() > 0);
    const cv::Mat d = pMP->GetDescriptor();
    desc.insert(desc.end(), d.data, d.data + 32);
    src.push_back(i);
  }
  const int M = (int)src.size(), NL = CurrentFrame.Nleft;
  // ---- left camera: every projected point ----
  CameraFlat left(CurrentFrame, false);
  orbx_projected ptsL{M, u.data(), v.data(), nullptr, radius.data(), lo.data(), hi.data(), angle.data(), has_obs.data(),
                      desc.data()};
  std::vector<int32_t> decL(M, -1), winL(M, 0);
  if (orbm_search_by_projection_frame_decisions(OrbxThreadMatcher(), &left.v, &ptsL, TH_HIGH, decL.data(), winL.data()) !=
      ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  // ---- right camera: the points whose LEFT window held a feature (`if (vIndices2.empty()) continue;`, :1655) ----
  std::vector<int> keep;
  for (int k = 0; k < M; k++)
    if (winL[k] > 0) keep.push_back(k);
  const int MR = (int)keep.size();
  std::vector<float> uR(MR), vR(MR), radR(MR), angR(MR);
  std::vector<int32_t> loR(MR), hiR(MR);
  std::vector<uint8_t> obsR(MR), descR((size_t)MR * 32);
  for (int j = 0; j < MR; j++) {
    const int k = keep[j];
    uR[j] = ur2[k]; vR[j] = vr2[k]; radR[j] = radius[k]; angR[j] = angle[k];
    loR[j] = lo[k]; hiR[j] = hi[k]; obsR[j] = has_obs[k];
    memcpy(&descR[(size_t)j * 32], &desc[(size_t)k * 32], 32);
  }
  CameraFlat right(CurrentFrame, true);
  orbx_projected ptsR{MR, uR.data(), vR.data(), nullptr, radR.data(), loR.data(), hiR.data(), angR.data(), obsR.data(),
                      descR.data()};
  std::vector<int32_t> decR(MR, -1);
  if (orbm_search_by_projection_frame_decisions(OrbxThreadMatcher(), &right.v, &ptsR, TH_HIGH, decR.data(), nullptr) !=
      ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  // ---- the reference's bookkeeping in its own order: left then right of every point (:1692-1706, :1757-1778) ----
  int nmatches = 0;
  const float factor = 1.0f / HISTO_LENGTH;
  auto bin_of = [&](float a_last, float a_cur) {
    float rot = a_last - a_cur;
    if (rot < 0.0) rot += 360.0f;
    int bin = round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
  };
  for (int k = 0, j = 0; k < M; k++) {
    MapPoint* pMP = LastFrame.mvpMapPoints[src[k]];
    if (decL[k] >= 0) {
      CurrentFrame.mvpMapPoints[decL[k]] = pMP;
      nmatches++;
      if (mbCheckOrientation) rotHist[bin_of(angle[k], CurrentFrame.mvKeys[decL[k]].angle)].push_back(decL[k]);
    }
    if (j < MR && keep[j] == k) {
      if (decR[j] >= 0) {
        CurrentFrame.mvpMapPoints[decR[j] + NL] = pMP;
        nmatches++;
        if (mbCheckOrientation)
          rotHist[bin_of(angle[k], CurrentFrame.mvKeysRight[decR[j]].angle)].push_back(decR[j] + NL);
      }
      j++;
    }
  }
  return nmatches;
}
}  // namespace

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
  if (CurrentFrame.Nleft != -1) {
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = SearchByProjectionTwoCameraFrames(CurrentFrame, LastFrame, th, bMono, mbCheckOrientation, rotHist);
    if (mbCheckOrientation) {  // :1785-1803
      int ind1 = -1, ind2 = -1, ind3 = -1;
      ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
      for (int i = 0; i < HISTO_LENGTH; i++) {
        if (i != ind1 && i != ind2 && i != ind3) {
          for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
            CurrentFrame.mvpMapPoints[rotHist[i][j]] = static_cast<MapPoint*>(NULL);
            nmatches--;
          }
        }
      }
    }
    return nmatches;
  }
  // caller-side projection, exactly the reference's expressions (:1607-1690): Tcw, forward / backward, x3Dc, uv, level
  // window and radius are computed here on the host with Sophus / Eigen; only points that pass those tests are sent
  const Sophus::SE3f Tcw = CurrentFrame.GetPose();
  const Eigen::Vector3f twc = Tcw.inverse().translation();
  const Sophus::SE3f Tlw = LastFrame.GetPose();
  const Eigen::Vector3f tlc = Tlw * twc;
  const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;
  const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;
  std::vector<float> u, v, ur, radius, angle;
  std::vector<int32_t> lo, hi, src;
  std::vector<uint8_t> has_obs, desc;
  for (int i = 0; i < LastFrame.N; i++) {
    MapPoint* pMP = LastFrame.mvpMapPoints[i];
    if (!pMP || LastFrame.mvbOutlier[i]) continue;
    const Eigen::Vector3f x3Dc = Tcw * pMP->GetWorldPos();
    const float invzc = 1.0 / x3Dc(2);
    if (invzc < 0) continue;
    const Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
    if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
    if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
    const int nLastOctave = LastFrame.mvKeys[i].octave;
    u.push_back(uv(0)); v.push_back(uv(1));
    ur.push_back(uv(0) - CurrentFrame.mbf * invzc);
    radius.push_back(th * CurrentFrame.mvScaleFactors[nLastOctave]);
    lo.push_back(bForward ? nLastOctave : (bBackward ? 0 : nLastOctave - 1));
    hi.push_back(bForward ? -1 : (bBackward ? nLastOctave : nLastOctave + 1));
    angle.push_back(LastFrame.mvKeysUn[i].angle);
    has_obs.push_back(pMP->Observations() > 0);  // temporal stereo points of UpdateLastFrame have none: they do not block (:1663)
    const cv::Mat d = pMP->GetDescriptor();      // a clone per call: take the range from ONE of them
    desc.insert(desc.end(), d.data, d.data + 32);
    src.push_back(i);
  }
  FrameFlat ff(CurrentFrame);
  orbx_projected pts{(int32_t)u.size(), u.data(), v.data(), CurrentFrame.mvuRight.empty() ? nullptr : ur.data(),
                     radius.data(), lo.data(), hi.data(), angle.data(), has_obs.data(), desc.data()};
  std::vector<int32_t> assign(CurrentFrame.N, -1);
  int32_t nmatches = 0;
  if (orbm_search_by_projection_frame(OrbxThreadMatcher(), &ff.v, &pts, TH_HIGH, mbCheckOrientation, assign.data(),
                                      &nmatches) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  for (int i = 0; i < CurrentFrame.N; i++)
    if (assign[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[src[assign[i]]];
  return nmatches;
}

namespace {
// DBoW2::FeatureVector + MapPoint flags of a KeyFrame -> orbx_keyframe_view (kps filled in by the caller)
struct TriFlat {
  std::vector<uint8_t> has_mp;
  std::vector<uint32_t> ids, idx;
  std::vector<int32_t> off{0};
  orbx_keyframe_view v;
  explicit TriFlat(KeyFrame* kf) : has_mp(kf->N) {
    for (int i = 0; i < kf->N; i++) has_mp[i] = kf->GetMapPoint(i) != nullptr;
    for (const auto& node : kf->mFeatVec) {
      ids.push_back(node.first);
      idx.insert(idx.end(), node.second.begin(), node.second.end());
      off.push_back((int32_t)idx.size());
    }
    v.n = kf->N;
    v.kps = nullptr;
    v.desc = kf->mDescriptors.data;
    v.u_right = nullptr;
    v.has_mappoint = has_mp.data();
    v.featvec = orbx_featvec{(int32_t)ids.size(), ids.data(), off.data(), idx.data()};
    v.scale_factors = kf->mvScaleFactors.data();
    v.level_sigma2 = kf->mvLevelSigma2.data();
    v.n_levels = (int32_t)kf->mvScaleFactors.size();
  }
};

// Both KeyFrames of a two-camera rig (:913-921, :958-971, :991-994, :1007-1052). The descriptor part — every pair under
// a shared vocabulary node with distance <= TH_LOW, in scan order — comes from the device
// (orbm_triangulation_candidates); the epipolar test is KannalaBrandt8::epipolarConstrain (= TriangulateMatches), which
// stays the reference's own camera code, asked here candidate by candidate exactly where the reference asks it.
template <class ThreeMaxima>  // ORBmatcher::ComputeThreeMaxima is a protected member: handed in by the caller
int SearchForTriangulationTwoCameras(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t>>& vMatchedPairs,
                                     bool bOnlyStereo, bool bCoarse, bool bCheckOrientation, ThreeMaxima three_maxima) {
  const int TH_LOW = ORBmatcher::TH_LOW, HISTO_LENGTH = ORBmatcher::HISTO_LENGTH;
  vMatchedPairs.clear();
  if (bOnlyStereo) return 0;  // bStereo1 = (!mpCamera2 && ...) is false for every feature                   :958-960
  const Sophus::SE3f T1w = pKF1->GetPose(), Tw2 = pKF2->GetPoseInverse();                           // :896-921
  const Sophus::SE3f Tr1w = pKF1->GetRightPose(), Twr2 = pKF2->GetRightPoseInverse();
  const Sophus::SE3f T[4] = {T1w * Tw2, T1w * Twr2, Tr1w * Tw2, Tr1w * Twr2};                       // ll, lr, rl, rr
  Eigen::Matrix3f R[4];
  Eigen::Vector3f t[4];
  for (int k = 0; k < 4; k++) {
    R[k] = T[k].rotationMatrix();
    t[k] = T[k].translation();
  }
  TriFlat k1(pKF1), k2(pKF2);
  auto key = [](KeyFrame* kf, int i) -> const cv::KeyPoint& {                                       // :966-969
    return kf->NLeft == -1 ? kf->mvKeysUn[i] : (i < kf->NLeft ? kf->mvKeys[i] : kf->mvKeysRight[i - kf->NLeft]);
  };
  std::vector<cv::KeyPoint> keys1(pKF1->N), keys2(pKF2->N);
  for (int i = 0; i < pKF1->N; i++) keys1[i] = key(pKF1, i);
  for (int i = 0; i < pKF2->N; i++) keys2[i] = key(pKF2, i);
  k1.v.kps = reinterpret_cast<const orbx_kp*>(keys1.data());
  k2.v.kps = reinterpret_cast<const orbx_kp*>(keys2.data());
  std::vector<int32_t> off(pKF1->N + 1), ci((size_t)16 * pKF1->N + 64), cd(ci.size());
  int32_t total = 0;
  int rc = orbm_triangulation_candidates(OrbxThreadMatcher(), &k1.v, &k2.v, off.data(), ci.data(), cd.data(),
                                         (int32_t)ci.size(), &total);
  if (rc == ORBX_E_CAPACITY) {
    ci.resize(total);
    cd.resize(total);
    rc = orbm_triangulation_candidates(OrbxThreadMatcher(), &k1.v, &k2.v, off.data(), ci.data(), cd.data(), total, &total);
  }
  if (rc != ORBX_OK) throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  std::vector<int> vMatches12(pKF1->N, -1);
  std::vector<std::vector<int>> rotHist(HISTO_LENGTH);
  const float factor = 1.0f / HISTO_LENGTH;
  int nmatches = 0;
  // rows in the reference's order: vocabulary nodes ascending, features in the node's list order (only the rotation
  // histogram's push order depends on it, and that order decides nothing)
  for (int idx1 = 0; idx1 < pKF1->N; idx1++) {
    if (off[idx1] == off[idx1 + 1]) continue;
    const cv::KeyPoint& kp1 = keys1[idx1];
    const bool bRight1 = !(pKF1->NLeft == -1 || idx1 < pKF1->NLeft);
    int bestDist = TH_LOW, bestIdx2 = -1;
    for (int c = off[idx1]; c < off[idx1 + 1]; c++) {
      const int idx2 = ci[c], dist = cd[c];
      if (dist > bestDist) continue;                                                                // :988
      const cv::KeyPoint& kp2 = keys2[idx2];
      const bool bRight2 = !(pKF2->NLeft == -1 || idx2 < pKF2->NLeft);
      const int pair = 2 * (int)bRight1 + (int)bRight2;                                             // :1007-1043
      GeometricCamera* pCamera1 = bRight1 ? pKF1->mpCamera2 : pKF1->mpCamera;
      GeometricCamera* pCamera2 = bRight2 ? pKF2->mpCamera2 : pKF2->mpCamera;
      if (bCoarse || pCamera1->epipolarConstrain(pCamera2, kp1, kp2, R[pair], t[pair], pKF1->mvLevelSigma2[kp1.octave],
                                                 pKF2->mvLevelSigma2[kp2.octave])) {               // :1045-1052
        bestIdx2 = idx2;
        bestDist = dist;
      }
    }
    if (bestIdx2 >= 0) {                                                                            // :1060-1076
      vMatches12[idx1] = bestIdx2;
      nmatches++;
      if (bCheckOrientation) {
        float rot = kp1.angle - keys2[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(idx1);
      }
    }
  }
  if (bCheckOrientation) {                                                                          // :1082-1095
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist.data(), HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
        vMatches12[rotHist[i][j]] = -1;
        nmatches--;
      }
    }
  }
  vMatchedPairs.reserve(nmatches);
  for (size_t i = 0, iend = vMatches12.size(); i < iend; i++)
    if (vMatches12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)vMatches12[i]));
  return nmatches;
}
}  // namespace

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2,
                                       std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo,
                                       const bool bCoarse) {
  if (pKF1->mpCamera2 && pKF2->mpCamera2)
    return SearchForTriangulationTwoCameras(
        pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse, mbCheckOrientation,
        [this](std::vector<int>* h, int L, int& a, int& b, int& c) { ComputeThreeMaxima(h, L, a, b, c); });
  // epipole and F12 exactly as the reference computes them (:893-911, src/CameraModels/Pinhole.cpp:122-149)
  const Sophus::SE3f T1w = pKF1->GetPose(), T2w = pKF2->GetPose(), Tw2 = pKF2->GetPoseInverse();
  const Eigen::Vector3f C2 = T2w * pKF1->GetCameraCenter();
  const Eigen::Vector2f ep = pKF2->mpCamera->project(C2);
  const Sophus::SE3f T12 = T1w * Tw2;
  const Eigen::Matrix3f R12 = T12.rotationMatrix();
  const Eigen::Vector3f t12 = T12.translation();
  Eigen::Matrix3f t12x;
  t12x << 0, -t12(2), t12(1), t12(2), 0, -t12(0), -t12(1), t12(0), 0;
  const Eigen::Matrix3f K1 = pKF1->mpCamera->toK_(), K2 = pKF2->mpCamera->toK_();
  const Eigen::Matrix3f F12 = K1.transpose().inverse() * t12x * R12 * K2.inverse();
  float F[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) F[3 * r + c] = F12(r, c);

  auto flatten = [](KeyFrame* kf, std::vector<uint8_t>& has_mp, std::vector<uint32_t>& ids, std::vector<int32_t>& off,
                    std::vector<uint32_t>& idx, orbx_keyframe_view* out) {
    has_mp.resize(kf->N);
    for (int i = 0; i < kf->N; i++) has_mp[i] = kf->GetMapPoint(i) != nullptr;
    off.push_back(0);
    for (const auto& node : kf->mFeatVec) {  // DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>>
      ids.push_back(node.first);
      idx.insert(idx.end(), node.second.begin(), node.second.end());
      off.push_back((int32_t)idx.size());
    }
    out->n = kf->N;
    out->kps = reinterpret_cast<const orbx_kp*>(kf->mvKeysUn.data());
    out->desc = kf->mDescriptors.data;
    out->u_right = kf->mvuRight.empty() ? nullptr : kf->mvuRight.data();
    out->has_mappoint = has_mp.data();
    out->featvec = orbx_featvec{(int32_t)ids.size(), ids.data(), off.data(), idx.data()};
    out->scale_factors = kf->mvScaleFactors.data();
    out->level_sigma2 = kf->mvLevelSigma2.data();
    out->n_levels = (int32_t)kf->mvScaleFactors.size();
  };
  std::vector<uint8_t> hm1, hm2;
  std::vector<uint32_t> ids1, ids2, idx1, idx2;
  std::vector<int32_t> off1, off2;
  orbx_keyframe_view v1, v2;
  flatten(pKF1, hm1, ids1, off1, idx1, &v1);
  flatten(pKF2, hm2, ids2, off2, idx2, &v2);
  std::vector<int32_t> m12(pKF1->N, -1);
  int32_t nmatches = 0;
  if (orbm_search_for_triangulation(OrbxThreadMatcher(), &v1, &v2, F, ep(0), ep(1), bOnlyStereo, bCoarse,
                                    mbCheckOrientation, m12.data(), &nmatches) != ORBX_OK)
    throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
  vMatchedPairs.clear();  // :1097-1103, ascending idx1
  vMatchedPairs.reserve(nmatches);
  for (size_t i = 0; i < m12.size(); i++)
    if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
  return nmatches;
}

}  // namespace ORB_SLAM3
