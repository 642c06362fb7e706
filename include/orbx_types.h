/* orbx_types.h — plain-old-data types that cross the orbx C ABI.
 *
 * Every struct here is the flat ("SoA / CSR") form of an object the reference keeps as C++ containers; the
 * reference type each one replaces is cited beside it (paths relative to the reference checkout).
 * No C++ types, no torch types: a maintainer can bind these from C, C++, ctypes or cffi.
 */
#ifndef ORBX_TYPES_H_
#define ORBX_TYPES_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint, 28 bytes, standard layout {Point2f pt; float size, angle, response; int octave, class_id;}.
 * `std::vector<cv::KeyPoint>::data()` can be reinterpret_cast to orbx_kp* (include/Frame.h:254-255). */
typedef struct orbx_kp {
  float x, y;
  float size;
  float angle;    /* degrees, [0,360) */
  float response; /* FAST corner score (src/ORBextractor.cc:810-815) */
  int32_t octave;
  int32_t class_id; /* always -1 */
} orbx_kp;

#define ORBX_DESC_BYTES 32 /* one row of the N x 32 CV_8U descriptor cv::Mat (include/Frame.h:268) */

#define ORBX_GRID_COLS 64 /* FRAME_GRID_COLS include/Frame.h:42 */
#define ORBX_GRID_ROWS 48 /* FRAME_GRID_ROWS include/Frame.h:41 */

/* Frame::mGrid[64][48] (include/Frame.h:279) as CSR. Cell id = col * 48 + row (mGrid[col][row]); items inside a
 * cell are ascending keypoint indices, the order Frame::AssignFeaturesToGrid produces (src/Frame.cc:520-547). */
typedef struct orbx_grid {
  const int32_t* cell_offsets; /* [64*48 + 1] */
  const int32_t* cell_items;   /* [cell_offsets[64*48]] */
  float min_x, min_y;          /* Frame::mnMinX, mnMinY */
  float inv_w, inv_h;          /* Frame::mfGridElementWidthInv / HeightInv */
} orbx_grid;

/* The part of a Frame that the guided searches read (Nleft == -1, i.e. pinhole mono / stereo / RGB-D):
 * mvKeysUn, mDescriptors, mvuRight, mGrid, mvScaleFactors (include/Frame.h:254-279). */
typedef struct orbx_frame_view {
  int32_t n;
  const orbx_kp* kps;   /* mvKeysUn */
  const uint8_t* desc;  /* n x 32 */
  const float* u_right; /* mvuRight or NULL (monocular) */
  /* per keypoint: 1 if mvpMapPoints[i] != NULL && mvpMapPoints[i]->Observations() > 0 (src/ORBmatcher.cc:92-93) */
  const uint8_t* occupied;
  orbx_grid grid;
  const float* scale_factors; /* mvScaleFactors[n_levels] */
  int32_t n_levels;
} orbx_frame_view;

/* Local-map points as flattened by the shim from MapPoint tracking scratch (include/MapPoint.h:172-180), already
 * filtered through Frame::isInFrustum (src/Frame.cc:632-699). One entry per element of vpMapPoints. */
typedef struct orbx_mappoints {
  int32_t m;
  const uint8_t* track_in_view; /* mbTrackInView && !isBad() */
  const float* proj_x;          /* mTrackProjX */
  const float* proj_y;          /* mTrackProjY */
  const float* proj_xr;         /* mTrackProjXR */
  const int32_t* level;         /* mnTrackScaleLevel */
  const float* view_cos;        /* mTrackViewCos */
  const float* depth;           /* mTrackDepth */
  const uint8_t* has_obs;       /* Observations() > 0: a keypoint that receives this point blocks later ones */
  const uint8_t* desc;          /* m x 32, MapPoint::GetDescriptor() */
} orbx_mappoints;

/* A two-camera (fisheye rig) Frame as the guided searches read it (Nleft != -1, include/Frame.h:254-279, 345-357):
 * mvKeys (left, n_left) and mvKeysRight (n_right) — NOT undistorted copies, Frame::GetFeaturesInArea reads mvKeys /
 * mvKeysRight in this mode (src/Frame.cc:811-813) —, mDescriptors with the left rows first and the right rows after
 * them (N = Nleft + Nright), one occupancy flag per row of mvpMapPoints[N], mGrid / mGridRight (same bounds and cell
 * sizes, indices local to their camera), and the left <-> right stereo associations of
 * Frame::ComputeStereoFishEyeMatches (mvLeftToRightMatch[Nleft], mvRightToLeftMatch[Nright], -1 = none). */
typedef struct orbx_fisheye_view {
  int32_t n_left, n_right;
  const orbx_kp* kps_left;
  const orbx_kp* kps_right;
  const uint8_t* desc;          /* (n_left + n_right) x 32 */
  const uint8_t* occupied;      /* [n_left + n_right]: mvpMapPoints[i] != NULL && Observations() > 0 */
  orbx_grid grid_left;
  orbx_grid grid_right;         /* min_x / min_y / inv_w / inv_h are read from grid_left */
  const int32_t* left_to_right; /* [n_left] */
  const int32_t* right_to_left; /* [n_right] */
  const float* scale_factors;
  int32_t n_levels;
} orbx_fisheye_view;

/* The right-camera tracking scratch of the local-map points (include/MapPoint.h:172-180), one entry per element of the
 * orbx_mappoints it accompanies. */
typedef struct orbx_mappoints_right {
  const uint8_t* track_in_view_r; /* mbTrackInViewR && !isBad() */
  const float* proj_x_r;          /* mTrackProjXR */
  const float* proj_y_r;          /* mTrackProjYR */
  const int32_t* level_r;         /* mnTrackScaleLevelR (-1: the right search is skipped, src/ORBmatcher.cc:150) */
  const float* view_cos_r;        /* mTrackViewCosR */
} orbx_mappoints_right;

/* What bool Frame::isInFrustum(MapPoint* pMP, float viewingCosLimit) (src/Frame.cc:632-699, Nleft == -1) reads from
 * the Frame: the pose (mRcw row-major, mtcw, mOw; include/Frame.h:197-199), the pinhole intrinsics of mpCamera
 * (src/CameraModels/Pinhole.cpp:47-53), mbf, the image bounds (include/Frame.h:372-375) and the scale pyramid
 * (mfLogScaleFactor, mnScaleLevels; :305-307). 104 bytes, no padding; arrays of it cross the ABI for batched calls. */
typedef struct orbx_frustum {
  float Rcw[9];
  float tcw[3];
  float Ow[3];
  float fx, fy, cx, cy;
  float mbf;
  float min_x, max_x, min_y, max_y; /* mnMinX, mnMaxX, mnMinY, mnMaxY */
  float log_scale_factor;           /* mfLogScaleFactor = logf(mfScaleFactor) */
  int32_t n_levels;                 /* mnScaleLevels */
} orbx_frustum;

/* mvpLocalMapPoints as the loop of Tracking::SearchLocalPoints (src/Tracking.cc:3288-3300) reads it, before the
 * projection: one entry per MapPoint*. n_maps maps of m points each may be stored back to back ([n_maps][m] arrays);
 * the single-frame calls use n_maps = 1. */
typedef struct orbx_local_map {
  int32_t m;
  int32_t n_maps;
  const float* pos;       /* [m][3] MapPoint::GetWorldPos() */
  const float* normal;    /* [m][3] MapPoint::GetNormal() */
  const float* min_dist;  /* mfMinDistance (GetMinDistanceInvariance() = 0.8f * it, src/MapPoint.cc:533-541) */
  const float* max_dist;  /* mfMaxDistance (GetMaxDistanceInvariance() = 1.2f * it; PredictScale reads it raw, :559-573) */
  const uint8_t* skip;    /* mnLastFrameSeen == mCurrentFrame.mnId || isBad() (src/Tracking.cc:3289) */
  const uint8_t* has_obs; /* Observations() > 0 (src/ORBmatcher.cc:92-93) */
  const uint8_t* desc;    /* [m][32] MapPoint::GetDescriptor() */
} orbx_local_map;

/* Scalars of one Tracking::SearchLocalPoints pass (src/Tracking.cc:3288-3330). */
typedef struct orbx_track_params {
  float viewing_cos_limit; /* isInFrustum(pMP, 0.5) */
  float th;                /* SearchByProjection's th: 1, 3 (RGB-D), 2 / 6 / 10 (IMU states), 5, 15 (:3303-3322) */
  float nnratio;           /* ORBmatcher matcher(0.8) */
  int32_t far_points;      /* mpLocalMapper->mbFarPoints */
  float th_far;            /* mpLocalMapper->mThFarPoints */
  float min_x, min_y;      /* Frame::mnMinX, mnMinY */
  float inv_w, inv_h;      /* Frame::mfGridElementWidthInv / HeightInv */
  int32_t cand_per_frame;  /* capacity of the per-frame candidate list, counted in grid records: the keypoints of the
                              cells every searched point's window touches (an upper bound of its candidates);
                              0 = 16 * m * max(1, th^2 / 4), at most m * cap. A frame that needs more is reported
                              with status ORBX_E_CAPACITY. */
} orbx_track_params;

/* Points of the last frame / a keyframe already projected into the current frame by the caller (the SE3 product and
 * camera projection stay on the host so Eigen's evaluation order is untouched): src/ORBmatcher.cc:1617-1660,
 * 1826-1853. One entry per candidate point that survived the caller-side tests. */
typedef struct orbx_projected {
  int32_t m;
  const float* u;          /* projected column */
  const float* v;          /* projected row */
  const float* u_right;    /* u - bf * invz, used only when the frame has stereo (NULL otherwise) */
  const float* radius;     /* th * scale[level] */
  const int32_t* min_level;
  const int32_t* max_level;
  const float* angle;      /* orientation of the source keypoint, degrees */
  const uint8_t* has_obs;  /* does the point block the keypoint it is written to (see orbx_search_by_projection_frame) */
  const uint8_t* desc;     /* m x 32 */
} orbx_projected;

/* DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>, Thirdparty/DBoW2/DBoW2/FeatureVector.h) as CSR. */
typedef struct orbx_featvec {
  int32_t n_nodes;
  const uint32_t* node_ids; /* ascending */
  const int32_t* offsets;   /* [n_nodes + 1] */
  const uint32_t* indices;  /* keypoint indices, per node in insertion order */
} orbx_featvec;

/* The vocabulary tree of DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (Thirdparty/DBoW2/DBoW2/
 * TemplatedVocabulary.h: m_nodes with children / descriptor / word_id / weight, m_L) flattened node by node; node 0 is
 * the root. children of node n = children[child_offsets[n] .. child_offsets[n + 1]) in m_nodes[n].children order; a
 * node without children is a leaf (a word). */
typedef struct orbx_vocabulary {
  int32_t n_nodes;
  int32_t depth;                /* m_L */
  const int32_t* child_offsets; /* [n_nodes + 1] */
  const uint32_t* children;
  const uint8_t* descriptors;   /* n_nodes x 32 (the root's row is not read) */
  const uint32_t* word_id;      /* per node; meaningful for leaves */
  const double* weight;         /* per node; WordValue of the leaf */
} orbx_vocabulary;

/* KeyFrame view for SearchForTriangulation (pinhole, NLeft == -1): include/KeyFrame.h:384-401. */
typedef struct orbx_keyframe_view {
  int32_t n;
  const orbx_kp* kps;           /* mvKeysUn */
  const uint8_t* desc;          /* n x 32 */
  const float* u_right;         /* mvuRight (values < 0 = monocular point) or NULL */
  const uint8_t* has_mappoint;  /* GetMapPoint(i) != NULL */
  orbx_featvec featvec;
  const float* scale_factors;   /* mvScaleFactors */
  const float* level_sigma2;    /* mvLevelSigma2 */
  int32_t n_levels;
} orbx_keyframe_view;

#ifdef __cplusplus
}
#endif
#endif /* ORBX_TYPES_H_ */
