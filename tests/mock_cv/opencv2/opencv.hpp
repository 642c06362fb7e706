// Minimal stand-in for the OpenCV core types the ORBextractor shim touches — TEST INFRASTRUCTURE ONLY.
// The image has no OpenCV C++ headers; this mock lets tests compile shim/ORBextractor.cc and run it on the GPU box.
// Layouts that matter to the ABI (cv::KeyPoint = 28-byte POD, cv::Mat row-major with `step`) follow OpenCV's.
#ifndef MOCK_OPENCV_HPP_
#define MOCK_OPENCV_HPP_
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {
struct Point2f {
  float x, y;
};
struct Rect {
  int x, y, width, height;
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
class KeyPoint {
 public:
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
class Mat {
 public:
  int rows = 0, cols = 0;
  uint8_t* data = nullptr;
  size_t step = 0;
  Mat() {}
  Mat(int r, int c, int) { create(r, c, 0); }
  void create(int r, int c, int) {
    if (r == rows && c == cols && buf_) return;
    buf_.reset(new std::vector<uint8_t>((size_t)r * c));
    rows = r; cols = c; step = (size_t)c; data = buf_->data();
  }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  uint8_t* ptr(int r) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r) const { return data + (size_t)r * step; }
  Mat operator()(const Rect& r) const {
    Mat m;
    m.buf_ = buf_;
    m.rows = r.height; m.cols = r.width; m.step = step;
    m.data = data + (size_t)r.y * step + r.x;
    return m;
  }
 private:
  std::shared_ptr<std::vector<uint8_t>> buf_;
};
// _InputArray / _OutputArray: thin references to a Mat
class InputArray {
 public:
  InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
  InputArray() : m_(nullptr) {}
  bool empty() const { return !m_ || m_->empty(); }
  Mat getMat() const { return *m_; }
 protected:
  Mat* m_;
};
class OutputArray : public InputArray {
 public:
  OutputArray(Mat& m) : InputArray(m) {}
  void create(int r, int c, int t) const { m_->create(r, c, t); }
  void release() const { m_->release(); }
};
}  // namespace cv
#endif
