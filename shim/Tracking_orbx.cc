// shim/Tracking_orbx.cc — drop-in body of Tracking::SearchLocalPoints() (src/Tracking.cc:3249-3330), the caller of the
// local-map SearchByProjection: Frame::isInFrustum for the whole of mvpLocalMapPoints in ONE
// orbm_is_in_frustum call instead of one host call per MapPoint, then the drop-in
// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, ...) of shim/ORBmatcher_orbx.cc.
//
// COMPILES ONLY INSIDE THE REFERENCE TREE (needs include/Tracking.h and everything it pulls in). Replace the body of
// Tracking::SearchLocalPoints with the one below and add two accessors to include/MapPoint.h next to
// GetMinDistanceInvariance (mfMinDistance / mfMaxDistance are protected, :241-242, and MapPoint::PredictScale reads
// mfMaxDistance itself, src/MapPoint.cc:559-573, so 0.8f * / 1.2f * the accessor values cannot stand in for them):
//     float GetMinDistanceRaw() { unique_lock<mutex> lock(mMutexPos); return mfMinDistance; }
//     float GetMaxDistanceRaw() { unique_lock<mutex> lock(mMutexPos); return mfMaxDistance; }
// Every word the reference's loop leaves on a MapPoint (mbTrackInView, mTrackProjX / Y always; mTrackProjXR,
// mnTrackScaleLevel, mTrackViewCos, mTrackDepth only when the point is in view, src/Frame.cc:632-685), the visibility
// counters and mCurrentFrame.mmProjectPoints come out as the reference's (tests/test_shim_bodies_vs_reference_source.py:
// this body against the reference's own function text on the same stand-in Tracking object; on the GPU:
// tests/test_gpu_shim_bodies.py). Two-camera frames (Nleft != -1) keep the reference's per-point
// Frame::isInFrustum — its KannalaBrandt8 projection is camera-model code — and still take the drop-in
// two-camera SearchByProjection.
#include "Tracking.h"

#include <stdexcept>
#include <vector>

#include "ORBmatcher.h"
#include "orbm.h"
#include "orbx_thread_matcher.h"

namespace ORB_SLAM3 {

void Tracking::SearchLocalPoints() {
  Frame& F = mCurrentFrame;
  // :3268-3285 — the points the frame already holds: not searched again, seen in this frame. (The fork runs this loop
  // under tbb::parallel_for; every iteration touches its own slot and its own MapPoint, so the serial loop is the same.)
  for (MapPoint*& held : F.mvpMapPoints) {
    if (!held) continue;
    if (held->isBad()) {
      held = static_cast<MapPoint*>(nullptr);
      continue;
    }
    held->IncreaseVisible();
    held->mnLastFrameSeen = F.mnId;
    held->mbTrackInView = false;
    held->mbTrackInViewR = false;
  }

  // :3287-3300 — project the local map into the frame
  const int M = (int)mvpLocalMapPoints.size();
  int nToMatch = 0;
  if (F.Nleft != -1) {
    for (MapPoint* pMP : mvpLocalMapPoints) {
      if (pMP->mnLastFrameSeen == F.mnId || pMP->isBad()) continue;
      if (F.isInFrustum(pMP, 0.5)) {
        pMP->IncreaseVisible();
        nToMatch++;
      }
      if (pMP->mbTrackInView) F.mmProjectPoints[pMP->mnId] = cv::Point2f(pMP->mTrackProjX, pMP->mTrackProjY);
    }
  } else if (M > 0) {
    std::vector<float> pos((size_t)M * 3), normal((size_t)M * 3), dmin(M), dmax(M);
    std::vector<uint8_t> skip(M), in_view(M, 0);
    std::vector<float> px(M), py(M), pxr(M), vcos(M), depth(M);
    std::vector<int32_t> level(M);
    for (int i = 0; i < M; i++) {
      MapPoint* p = mvpLocalMapPoints[i];
      skip[i] = p->mnLastFrameSeen == F.mnId || p->isBad();  // :3289
      const Eigen::Vector3f P = p->GetWorldPos(), Pn = p->GetNormal();
      for (int k = 0; k < 3; k++) {
        pos[(size_t)i * 3 + k] = P(k);
        normal[(size_t)i * 3 + k] = Pn(k);
      }
      dmin[i] = p->GetMinDistanceRaw();
      dmax[i] = p->GetMaxDistanceRaw();
      // the fields the reference leaves alone on a point that is not in view make the round trip through the call
      px[i] = p->mTrackProjX; py[i] = p->mTrackProjY; pxr[i] = p->mTrackProjXR;
      level[i] = p->mnTrackScaleLevel; vcos[i] = p->mTrackViewCos; depth[i] = p->mTrackDepth;
    }
    orbx_frustum fr;
    const Sophus::SE3f Tcw = F.GetPose();  // mRcw / mtcw are its rotation / translation (Frame::UpdatePoseMatrices)
    const Eigen::Matrix3f Rcw = Tcw.rotationMatrix();
    const Eigen::Vector3f tcw = Tcw.translation(), Ow = F.GetOw();
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) fr.Rcw[3 * r + c] = Rcw(r, c);
      fr.tcw[r] = tcw(r);
      fr.Ow[r] = Ow(r);
    }
    fr.fx = F.mpCamera->getParameter(0);  // Pinhole::project (src/CameraModels/Pinhole.cpp:47-53)
    fr.fy = F.mpCamera->getParameter(1);
    fr.cx = F.mpCamera->getParameter(2);
    fr.cy = F.mpCamera->getParameter(3);
    fr.mbf = F.mbf;
    fr.min_x = Frame::mnMinX; fr.max_x = Frame::mnMaxX; fr.min_y = Frame::mnMinY; fr.max_y = Frame::mnMaxY;
    fr.log_scale_factor = F.mfLogScaleFactor;
    fr.n_levels = F.mnScaleLevels;
    orbx_local_map map{M, 1, pos.data(), normal.data(), dmin.data(), dmax.data(), skip.data(), nullptr, nullptr};
    int32_t n_in_view = 0;
    if (orbm_is_in_frustum(OrbxThreadMatcher(), &fr, &map, 0, 0.5f, in_view.data(), px.data(), py.data(), pxr.data(),
                           level.data(), vcos.data(), depth.data(), &n_in_view) != ORBX_OK)
      throw std::runtime_error(orbm_last_error(OrbxThreadMatcher()));
    for (int i = 0; i < M; i++) {
      if (skip[i]) continue;
      MapPoint* p = mvpLocalMapPoints[i];
      p->mbTrackInView = in_view[i] != 0;
      p->mTrackProjX = px[i];  // -1 unless the point projects into the image (src/Frame.cc:635-636, :655-656)
      p->mTrackProjY = py[i];
      if (!in_view[i]) continue;
      p->mTrackProjXR = pxr[i];  // src/Frame.cc:676-685
      p->mTrackDepth = depth[i];
      p->mnTrackScaleLevel = level[i];
      p->mTrackViewCos = vcos[i];
      p->IncreaseVisible();
      F.mmProjectPoints[p->mnId] = cv::Point2f(px[i], py[i]);
    }
    nToMatch = n_in_view;
  }

  if (nToMatch <= 0) return;
  // :3302-3322 — the search radius of the tracker's state
  const bool rgbd = mSensor == System::RGBD || mSensor == System::IMU_RGBD;
  const bool inertial = mSensor == System::IMU_MONOCULAR || mSensor == System::IMU_STEREO || mSensor == System::IMU_RGBD;
  int th = rgbd ? 3 : 1;
  if (mpAtlas->isImuInitialized())
    th = mpAtlas->GetCurrentMap()->GetIniertialBA2() ? 2 : 6;
  else if (inertial)
    th = 10;
  if (F.mnId < mnLastRelocFrameId + 2) th = 5;
  if (mState == LOST || mState == RECENTLY_LOST) th = 15;
  ORBmatcher matcher(0.8);
  matcher.SearchByProjection(F, mvpLocalMapPoints, th, mpLocalMapper->mbFarPoints, mpLocalMapper->mThFarPoints);
}

}  // namespace ORB_SLAM3
