// Minimal stand-in for the OpenCV C++ API that /root/reference/src/ORBextractor.cc uses, so that the reference's OWN
// extractor source compiles here without OpenCV headers or libraries (oracle/Makefile target _ref). TEST INFRASTRUCTURE:
// it exists only to check the oracle's restatement of the orchestration (quadtree, cell loop, level order, assembly,
// orientation, descriptor sampling) against the reference code itself.
//
// The image primitives are NOT OpenCV's code: cv::resize / GaussianBlur / FAST / copyMakeBorder / fastAtan2 forward to
// the oracle's restatements in orbref.cpp, which tests/test_oracle_primitives.py pins byte for byte to the real OpenCV
// 4.13 kernels through cv2. Types (Mat with ROI views, KeyPoint, Point_, Size, Rect, Input/OutputArray) carry just the
// members the reference touches, with OpenCV's semantics (default KeyPoint fields, Point *= float, create() on a
// matching size keeps the buffer, cvRound = round-half-even).
#ifndef ORBREF_STUB_OPENCV_HPP_
#define ORBREF_STUB_OPENCV_HPP_
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iterator>
#include <memory>
#include <string>
#include <sstream>
#include <fstream>
#include <iostream>
#include <vector>

#include "orbref.h"

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5 /* named by FORB::toMat32F, which is never called here */

static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <typename T>
static inline Point_<T>& operator*=(Point_<T>& a, float b) {
  a.x = (T)(a.x * b);
  a.y = (T)(a.y * b);
  return a;
}
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// 8-bit single-channel matrix header over a shared buffer; row / column ranges are views like OpenCV's
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar* data = nullptr;
  std::shared_ptr<std::vector<uchar>> buf;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(int r, int c, int, void* ext, size_t ext_step) : rows(r), cols(c), step(ext_step), data((uchar*)ext) {}
  void create(int r, int c, int) {
    if (data && r == rows && c == cols) return;  // OpenCV: a matching header keeps its buffer
    rows = r;
    cols = c;
    step = (size_t)c;
    buf = std::make_shared<std::vector<uchar>>((size_t)r * c + 1);
    data = buf->data();
  }
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    memset(m.data, 0, (size_t)r * c);
    return m;
  }
  void release() { *this = Mat(); }
  int type() const { return CV_8UC1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t step1() const { return step; }
  template <typename T>
  T& at(int y, int x) { return data[(size_t)y * step + x]; }
  template <typename T>
  const T& at(int y, int x) const { return data[(size_t)y * step + x]; }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T>
  T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
  template <typename T>
  const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
  Mat rowRange(int a, int b) const {
    Mat m = *this;
    m.data = data + (size_t)a * step;
    m.rows = b - a;
    return m;
  }
  Mat colRange(int a, int b) const {
    Mat m = *this;
    m.data = data + a;
    m.cols = b - a;
    return m;
  }
  Mat row(int i) const { return rowRange(i, i + 1); }
  Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols);
    return m;
  }
  void copyTo(Mat dst) const {  // dst is a header of the same size over existing memory (row views in the reference)
    assert(dst.rows == rows && dst.cols == cols);
    for (int y = 0; y < rows; y++) memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols);
  }
};

class _InputArray {
 public:
  const Mat* m;
  _InputArray(const Mat& mm) : m(&mm) {}
  bool empty() const { return m->empty(); }
  Mat getMat() const { return *m; }
};
class _OutputArray {
 public:
  Mat* m;
  _OutputArray(Mat& mm) : m(&mm) {}
  void create(int r, int c, int t) const { m->create(r, c, t); }
  void release() const { *m = Mat(); }
  Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

enum { NORM_L1 = 2 };
// cv::norm(a, b, NORM_L1) of two 8-bit views of one size: the sum of absolute differences (exact in a double)
static inline double norm(InputArray a, InputArray b, int type) {
  assert(type == NORM_L1);
  const Mat x = a.getMat(), y = b.getMat();
  assert(x.rows == y.rows && x.cols == y.cols);
  long long acc = 0;
  for (int r = 0; r < x.rows; r++)
    for (int c = 0; c < x.cols; c++) acc += std::abs((int)x.data[(size_t)r * x.step + c] - (int)y.data[(size_t)r * y.step + c]);
  return (double)acc;
}

static inline float fastAtan2(float y, float x) { return orbref_fast_atan2(y, x); }

static inline void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmax = true) {
  assert(nonmax);
  const Mat im = image.getMat();
  const int cap = im.rows * im.cols;
  std::vector<int> xs(cap + 1), ys(cap + 1), sc(cap + 1);
  const int n = orbref_fast9(im.data, im.cols, im.rows, (int)im.step, threshold, xs.data(), ys.data(), sc.data(), cap);
  keypoints.clear();
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

static inline void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sx, double sy, int border) {
  assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && border == BORDER_REFLECT_101);
  const Mat s = src.getMat().clone();  // the reference blurs in place
  dst.create(s.rows, s.cols, CV_8UC1);
  Mat d = dst.getMat();
  orbref_gauss7(s.data, s.cols, s.rows, (int)s.step, d.data, (int)d.step);
}

static inline void resize(InputArray src, OutputArray dst, Size sz, double, double, int interp) {
  assert(interp == INTER_LINEAR);
  const Mat s = src.getMat();
  dst.create(sz.height, sz.width, CV_8UC1);
  Mat d = dst.getMat();
  orbref_resize_linear(s.data, s.cols, s.rows, (int)s.step, d.data, d.cols, d.rows, (int)d.step);
}

static inline void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int type) {
  assert(top == bottom && top == left && top == right && (type & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  const Mat s = src.getMat().clone();  // the reference borders a level inside its own padded buffer
  dst.create(s.rows + 2 * top, s.cols + 2 * top, CV_8UC1);
  Mat d = dst.getMat();
  orbref_border101(s.data, s.cols, s.rows, (int)s.step, d.data, (int)d.step, top);
}


// cv::FileStorage / FileNode: DBoW2's TemplatedVocabulary has virtual save() / load() members that name them; they are
// never called here (the vocabulary is loaded with the reference's own loadFromTextFile), so these only have to parse.
struct FileNode {
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  template <typename T>
  operator T() const { return T(); }
  size_t size() const { return 0; }
};
// cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, matches, 2), the call of Frame::ComputeStereoFishEyeMatches
// (src/Frame.cc:1293). Defined in oracle/ref_matcher_shim.cpp over the oracle's knn2, which the primitive tests pin to
// the real cv2.BFMatcher.
struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0;
};
enum { NORM_HAMMING = 6 };
class BFMatcher {
 public:
  BFMatcher(int = NORM_HAMMING, bool = false) {}
  void knnMatch(const Mat& query, const Mat& train, std::vector<std::vector<DMatch>>& matches, int k) const;
};
struct FileStorage {
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
};
template <typename T>
static inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
#endif
