// orbx_plan.h — per-configuration geometry (host side, computed once per (w, h, extractor parameters)).
//
// Restates the set-up arithmetic of ORBextractor::ORBextractor (src/ORBextractor.cc:408-469), the level sizes of
// ComputePyramid (:1108-1118), the cell grid of ComputeKeyPointsOctTree (:768-784) and the root layout of
// DistributeOctTree (:566-585). The result is a POD that is passed to every kernel by value (__grid_constant__).
#ifndef ORBX_PLAN_H_
#define ORBX_PLAN_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "orbx_math.h"

namespace orbx {

constexpr int kMaxDim = 4096;  // candidates are packed x:12 | y:12 | score:8

struct LevelPlan {
  int w, h;           // cvRound((float)cols * mvInvScaleFactor[l])                     :1112-1113
  int pitch;          // row pitch (bytes) of the owned raw / blurred level buffers
  int64_t img_off;    // byte offset of the level inside one frame's pyramid slab
  int maxBX, maxBY;   // maxBorderX/Y = w - 16, h - 16 (minBorder = 16)                   :768-771
  int nCols, nRows;   // int(width / 35), int(height / 35)                                :781-782
  int wCell, hCell;   // ceil(width / nCols), ceil(height / nRows)                        :783-784
  int cell_base;      // first cell of this level in the per-frame cell numbering
  int slot_cap;       // candidate slots per cell = ceil(wCell/2) * ceil(hCell/2) (NMS packing bound)
  int slot_base;      // first slot (u32 units) of this level in the per-frame slot slab
  int quota;          // mnFeaturesPerLevel[l]                                            :436-448
  int nIni;           // round(width / height)                                            :566
  float hX;           // width / nIni                                                     :568
  int kp_cap;         // capacity of the per-level keypoint list (quota + overshoot)
  int kp_base;        // first entry of this level in the per-frame level-keypoint slab
  float scale, inv_scale, sigma2, inv_sigma2;  //                                        :418-432
  int patch;          // int(PATCH_SIZE * mvScaleFactor[l])                               :863
  int xtab_off, ytab_off;  // first entry of this level's resize tables (level >= 1)
};

struct Plan {
  int nlevels, w, h;
  int cells_per_frame, slots_per_frame, kps_per_frame;
  int64_t pyr_bytes_per_frame;
  int tab_entries;
  int max_tile_bytes;   // raw tile bytes for the largest FAST cell (pitch-padded)
  int max_cell_px;      // interior pixels of the largest cell
  int max_quota;
  int umax[16];         //                                                               :456-468
  LevelPlan lv[kMaxLevels];
};

// One axis of cv::resize(u8, INTER_LINEAR) (OpenCV imgproc/resize.cpp, INTER_RESIZE_COEF_BITS = 11), call site
// src/ORBextractor.cc:1122. clamp = the x axis behaviour (offsets clamped, fraction zeroed); rows are clipped at use.
inline void axis_table(int ssize, int dsize, bool clamp, int16_t* ofs, int16_t* c0, int16_t* c1) {
  const double inv_scale = (double)dsize / ssize;
  const double scale = 1. / inv_scale;
  for (int d = 0; d < dsize; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floor(f);
    f -= s;
    if (clamp) {
      if (s < 0) { f = 0; s = 0; }
      if (s >= ssize - 1) { f = 0; s = ssize - 1; }
    }
    int a = cv_round((1.f - f) * 2048.f), b = cv_round(f * 2048.f);
    a = a < -32768 ? -32768 : (a > 32767 ? 32767 : a);
    b = b < -32768 ? -32768 : (b > 32767 ? 32767 : b);
    ofs[d] = (int16_t)s;
    c0[d] = (int16_t)a;
    c1[d] = (int16_t)b;
  }
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// The image-size independent tables of the constructor (src/ORBextractor.cc:418-448): mvScaleFactor, mvLevelSigma2 and
// mnFeaturesPerLevel. Valid for every constructor-legal (scaleFactor, nlevels).
inline void make_tables(int nfeatures, float scale_factor_f, int nlevels, float* sf, float* s2, int* quota) {
  const double scaleFactor = scale_factor_f;  // include/ORBextractor.h:106 — a double member set from a float
  sf[0] = 1.0f;
  s2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    sf[i] = (float)(sf[i - 1] * scaleFactor);
    s2[i] = sf[i] * sf[i];
  }
  const float factor = (float)(1.0f / scaleFactor);
  float want = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; l++) {
    quota[l] = cv_round(want);
    sum += quota[l];
    want *= factor;
  }
  quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

// Returns 0, or a negative error: -1 bad arguments, -2 image too small for the level count, -3 too large.
inline int make_plan(int w, int h, int nfeatures, float scale_factor_f, int nlevels, Plan* P) {
  memset(P, 0, sizeof(*P));
  if (w <= 0 || h <= 0 || nlevels < 1 || nlevels > kMaxLevels || nfeatures < 1) return -1;
  if (w > kMaxDim || h > kMaxDim) return -3;
  P->nlevels = nlevels;
  P->w = w;
  P->h = h;
  float sf[kMaxLevels], s2[kMaxLevels];
  int quota[kMaxLevels];
  make_tables(nfeatures, scale_factor_f, nlevels, sf, s2, quota);
  {  // umax :456-468
    int v, v0;
    const int vmax = (int)floorf(kHalfPatch * sqrtf(2.f) / 2 + 1);
    const int vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v <= vmax; ++v) P->umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (P->umax[v0] == P->umax[v0 + 1]) ++v0;
      P->umax[v] = v0;
      ++v0;
    }
  }
  int64_t off = 0;
  int cells = 0, slots = 0, kps = 0, tabs = 0;
  for (int l = 0; l < nlevels; l++) {
    LevelPlan& L = P->lv[l];
    L.scale = sf[l];
    L.inv_scale = 1.0f / sf[l];
    L.sigma2 = s2[l];
    L.inv_sigma2 = 1.0f / s2[l];
    L.w = cv_round((float)w * L.inv_scale);
    L.h = cv_round((float)h * L.inv_scale);
    L.pitch = round_up(L.w, 64);
    L.img_off = off;
    off += (int64_t)L.pitch * L.h;
    L.maxBX = L.w - kMinBorder;
    L.maxBY = L.h - kMinBorder;
    const float width = (float)(L.maxBX - kMinBorder), height = (float)(L.maxBY - kMinBorder);
    L.nCols = (int)(width / kCellW);
    L.nRows = (int)(height / kCellW);
    if (L.nCols < 1 || L.nRows < 1) return -2;  // the reference divides by zero here
    L.wCell = (int)ceilf(width / L.nCols);
    L.hCell = (int)ceilf(height / L.nRows);
    L.cell_base = cells;
    cells += L.nCols * L.nRows;
    L.slot_cap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
    L.slot_base = slots;
    slots += L.nCols * L.nRows * L.slot_cap;
    L.quota = quota[l];
    L.nIni = (int)roundf(width / height);
    if (L.nIni < 1) return -2;  // taller than 2:1 — the reference divides by zero (hX = width / 0)
    L.hX = width / L.nIni;
    const int overshoot = 4 * L.nIni > L.quota ? 4 * L.nIni : L.quota;
    L.kp_cap = overshoot + 4;
    L.kp_base = kps;
    kps += L.kp_cap;
    L.patch = (int)(kPatchSize * sf[l]);
    L.xtab_off = tabs;
    tabs += l ? L.w : 0;
    L.ytab_off = tabs;
    tabs += l ? L.h : 0;
    const int tile = round_up(L.wCell + 6 + 4, 4) * (L.hCell + 6);
    if (tile > P->max_tile_bytes) P->max_tile_bytes = tile;
    if (L.wCell * L.hCell > P->max_cell_px) P->max_cell_px = L.wCell * L.hCell;
    if (L.quota > P->max_quota) P->max_quota = L.quota;
  }
  P->cells_per_frame = cells;
  P->slots_per_frame = slots;
  P->kps_per_frame = kps;
  P->pyr_bytes_per_frame = off;
  P->tab_entries = tabs;
  return 0;
}

}  // namespace orbx

#endif  // ORBX_PLAN_H_
