// shim/cv_standin.h -- the few cv:: types the shim headers touch, ONLY for compiling the shim in this repo where
// OpenCV C++ is absent (syntax / ABI check in tests/test_abi_cpu.py).  A real integration uses OpenCV's own headers.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {
struct Rect { int x, y, width, height; Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {} };
struct KeyPoint {
  struct { float x, y; } pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt{0, 0}, size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float s, float a, float r, int o, int c) : pt{x, y}, size(s), angle(a), response(r), octave(o), class_id(c) {}
};
class Mat {
 public:
  int rows = 0, cols = 0, type_ = 0;
  size_t step = 0;
  uint8_t* data = nullptr;
  std::shared_ptr<std::vector<uint8_t>> buf;
  Mat() {}
  Mat(int r, int c, int t) { create(r, c, t); }
  void create(int r, int c, int t) {
    rows = r; cols = c; type_ = t;
    const size_t es = (t == CV_32F) ? 4 : 1;
    step = (size_t)c * es;
    buf = std::make_shared<std::vector<uint8_t>>((size_t)r * step);
    data = buf->data();
  }
  int type() const { return type_; }
  bool empty() const { return rows == 0 || cols == 0; }
  void release() { rows = cols = 0; data = nullptr; buf.reset(); }
  uint8_t* ptr(int r) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r) const { return data + (size_t)r * step; }
  template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  Mat operator()(const Rect& r) const {
    Mat m = *this;
    m.rows = r.height; m.cols = r.width;
    m.data = data + (size_t)r.y * step + r.x;
    return m;
  }
  Mat getMat() const { return *this; }
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace cv
