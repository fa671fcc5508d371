// oracle/standin/cv_standin.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// A minimal stand-in for the part of the OpenCV C++ API that the reference's own sources on the hot path use
// (src/ORBextractor.cc, src/ORBmatcher.cc, src/Frame.cc, src/KeyFrame.cc, src/MapPoint.cc, src/Map.cc), so that
// those files can be compiled UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/ (see
// oracle/Makefile, target `ref`).  OpenCV itself is not in this image; its C++ library cannot be linked.
//
// What is restated here is third-party arithmetic only, each piece following OpenCV's published behaviour:
//   * image primitives  cv::resize / copyMakeBorder / GaussianBlur / FAST / fastAtan2 / cvRound  -> cv_prims.hpp,
//     every one of which tests/test_oracle_cpu.py compares bit for bit with the real cv2 4.13.0 wheel;
//   * cv::Mat as a reference-counted dense matrix with ROI views;
//   * the float matrix expressions the reference writes (A*B, A*B+C, -A.t()*B, A+B, A-B, s*A, A/s), evaluated the way
//     OpenCV's MatExpr folds them into ONE gemm call: the len<=4 / flags==0 inline path accumulates a row in float,
//     left to right, and adds alpha/beta terms in double; everything else accumulates in double (GEMMSingleMul<float,double>);
//   * cv::norm (L2, double accumulation), Mat::dot (double accumulation), cv::undistortPoints (5 fixed-point iterations).
// Nothing of the reference's control flow lives here.
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;
typedef unsigned short ushort;

inline int cvRound(double v) { return (int)lrint(v); }     // round half to even
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

#include "cv_prims.hpp"

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}   // Point2i(float,float) truncation is
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }   // done by the caller's conversion
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
  Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
  Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };
template <typename T> struct Scalar_ { T val[4]; Scalar_(T a = 0, T b = 0, T c = 0, T d = 0) : val{a, b, c, d} {} };
typedef Scalar_<double> Scalar;

struct KeyPoint {   // 28 bytes, the field order of cv::KeyPoint
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8U }; };
template <> struct DataType<ushort> { enum { type = CV_16U }; };
template <> struct DataType<int> { enum { type = CV_32S }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };

inline size_t elem_size1(int type) {
  static const size_t s[8] = {1, 1, 2, 2, 4, 4, 8, 2};
  return s[CV_MAT_DEPTH(type)];
}

struct MatStep {
  size_t v;
  size_t esz1;
  MatStep() : v(0), esz1(1) {}
  operator size_t() const { return v; }
  size_t operator[](int i) const { return i == 0 ? v : esz1; }
};

class Mat;
struct MatInit {   // Mat::zeros / ones / eye : an initializer expression; assigning it to a Mat of the same shape and
  int rows, cols, type, kind;   // type fills that Mat IN PLACE (OpenCV: MatOp_Initializer::assign -> m.create is a no-op)
  operator Mat() const;
};
template <typename T> class Mat_;
template <typename T> struct MatCommaInit_;

class Mat {
 public:
  int flags, rows, cols;
  uchar* data;
  MatStep step;
  std::shared_ptr<std::vector<uchar>> buf;

  Mat() : flags(0), rows(0), cols(0), data(nullptr) {}
  Mat(int r, int c, int type) : flags(0), rows(0), cols(0), data(nullptr) { create(r, c, type); }
  Mat(Size sz, int type) : flags(0), rows(0), cols(0), data(nullptr) { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t step_ = 0) : flags(type), rows(r), cols(c), data((uchar*)ext) {
    step.esz1 = elem_size1(type);
    step.v = step_ ? step_ : (size_t)c * elemSize();
  }
  template <typename T> explicit Mat(const std::vector<T>& v) : flags(0), rows(0), cols(0), data(nullptr) {
    create((int)v.size(), 1, DataType<T>::type);
    if (!v.empty()) memcpy(data, v.data(), v.size() * sizeof(T));
  }
  Mat(const MatInit& e) : flags(0), rows(0), cols(0), data(nullptr) { *this = e; }
  Mat& operator=(const MatInit& e) {
    create(e.rows, e.cols, e.type);
    for (int i = 0; i < rows; ++i) memset(data + (size_t)i * step, 0, (size_t)cols * elemSize());
    if (e.kind) {
      for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j)
          if (e.kind == 1 || i == j) set_d(i, j, 1.0);
    }
    return *this;
  }

  int type() const { return flags & 0xFFF; }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize() const { return elem_size1(flags) * channels(); }
  size_t elemSize1() const { return elem_size1(flags); }
  size_t step1() const { return step.v / elemSize1(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * cols; }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return rows <= 1 || step.v == (size_t)cols * elemSize(); }

  void create(int r, int c, int type) {
    type &= 0xFFF;
    if (data && rows == r && cols == c && this->type() == type) return;
    flags = type; rows = r; cols = c;
    step.esz1 = elem_size1(type);
    step.v = (size_t)c * elemSize();
    buf = std::make_shared<std::vector<uchar>>((size_t)r * step.v + 64);
    data = (r > 0 && c > 0) ? buf->data() : nullptr;
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { buf.reset(); data = nullptr; rows = cols = 0; }

  static MatInit zeros(int r, int c, int type) { return MatInit{r, c, type, 0}; }
  static MatInit ones(int r, int c, int type) { return MatInit{r, c, type, 1}; }
  static MatInit eye(int r, int c, int type) { return MatInit{r, c, type, 2}; }

  uchar* ptr(int i = 0) { return data + (size_t)i * step.v; }
  const uchar* ptr(int i = 0) const { return data + (size_t)i * step.v; }
  template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step.v); }
  template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step.v); }
  template <typename T> T& at(int i, int j) { return ((T*)(data + (size_t)i * step.v))[j]; }
  template <typename T> const T& at(int i, int j) const { return ((const T*)(data + (size_t)i * step.v))[j]; }
  template <typename T> T& at(int i0) {   // OpenCV Mat::at(int): row vector / continuous -> linear, column vector -> row
    if (isContinuous() || rows == 1) return ((T*)data)[i0];
    if (cols == 1) return *(T*)(data + (size_t)i0 * step.v);
    int i = i0 / cols, j = i0 - i * cols;
    return ((T*)(data + (size_t)i * step.v))[j];
  }
  template <typename T> const T& at(int i0) const { return const_cast<Mat*>(this)->at<T>(i0); }

  Mat rowRange(int a, int b) const { Mat m(*this); m.rows = b - a; m.data = data + (size_t)a * step.v; return m; }
  Mat colRange(int a, int b) const { Mat m(*this); m.cols = b - a; m.data = data + (size_t)a * elemSize(); return m; }
  Mat row(int i) const { return rowRange(i, i + 1); }
  Mat col(int j) const { return colRange(j, j + 1); }
  Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
  Mat operator()(Range rr, Range cr) const { return rowRange(rr.start, rr.end).colRange(cr.start, cr.end); }

  Mat clone() const { Mat m; copyTo(m); return m; }
  void copyTo(Mat& dst) const {
    if (empty()) { dst.release(); return; }
    dst.create(rows, cols, type());
    for (int i = 0; i < rows; ++i) memmove(dst.ptr(i), ptr(i), (size_t)cols * elemSize());
  }
  void copyTo(Mat&& dst) const { Mat d(dst); copyTo(d); }   // Rwc.copyTo(Twc.rowRange(0,3).colRange(0,3))
  Mat reshape(int cn, int newRows = 0) const {
    (void)newRows;
    Mat m(*this);
    assert(isContinuous());
    size_t rowElems = (size_t)cols * channels();
    assert(rowElems % cn == 0);
    m.flags = CV_MAKETYPE(depth(), cn);
    m.cols = (int)(rowElems / cn);
    return m;
  }
  void convertTo(Mat& dst, int rtype) const {
    Mat src = (dst.data == data) ? clone() : *this;
    dst.create(rows, cols, CV_MAKETYPE(CV_MAT_DEPTH(rtype), channels()));
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols * channels(); ++j) dst.set_d(i, j, src.get_d(i, j));
  }
  void push_back(const Mat& m) {
    Mat out(rows + m.rows, empty() ? m.cols : cols, empty() ? m.type() : type());
    for (int i = 0; i < rows; ++i) memcpy(out.ptr(i), ptr(i), (size_t)cols * elemSize());
    for (int i = 0; i < m.rows; ++i) memcpy(out.ptr(rows + i), m.ptr(i), (size_t)m.cols * m.elemSize());
    *this = out;
  }
  Mat& setTo(const Scalar& s) {
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols * channels(); ++j) set_d(i, j, s.val[0]);
    return *this;
  }
  Mat& operator=(const Scalar& s) { return setTo(s); }

  double get_d(int i, int j) const {   // j counts scalars (channels interleaved)
    switch (depth()) {
      case CV_8U: return ptr<uchar>(i)[j];
      case CV_16U: return ptr<ushort>(i)[j];
      case CV_32S: return ptr<int>(i)[j];
      case CV_32F: return ptr<float>(i)[j];
      case CV_64F: return ptr<double>(i)[j];
    }
    throw std::runtime_error("cv_standin: depth");
  }
  void set_d(int i, int j, double v) {
    switch (depth()) {
      case CV_8U: ptr<uchar>(i)[j] = (uchar)std::min(255, std::max(0, cvRound(v))); return;
      case CV_16U: ptr<ushort>(i)[j] = (ushort)std::min(65535, std::max(0, cvRound(v))); return;
      case CV_32S: ptr<int>(i)[j] = cvRound(v); return;
      case CV_32F: ptr<float>(i)[j] = (float)v; return;
      case CV_64F: ptr<double>(i)[j] = v; return;
    }
    throw std::runtime_error("cv_standin: depth");
  }

  struct TExpr;
  TExpr t() const;
  double dot(const Mat& m) const {   // dotProd_<T>: double accumulation of (double)a*b
    double r = 0;
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols * channels(); ++j) r += get_d(i, j) * m.get_d(i, j);
    return r;
  }
  Mat inv() const;
};

inline MatInit::operator Mat() const { Mat m; m = *this; return m; }

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  Mat_(const Mat& m) : Mat(m) {}
  T& operator()(int i, int j) { return this->template at<T>(i, j); }
  const T& operator()(int i, int j) const { return this->template at<T>(i, j); }
};
template <typename T> struct MatCommaInit_ {
  Mat_<T> m;
  int idx;
  MatCommaInit_(const Mat_<T>& m_, T v) : m(m_), idx(0) { put(v); }
  void put(T v) { m.template at<T>(idx / m.cols, idx % m.cols) = v; ++idx; }
  template <typename U> MatCommaInit_& operator,(U v) { put((T)v); return *this; }
  operator Mat() const { return m; }
  operator Mat_<T>() const { return m; }
};
template <typename T, typename U> MatCommaInit_<T> operator<<(const Mat_<T>& m, U v) { return MatCommaInit_<T>(m, (T)v); }

// ---------------------------------------------------------------------------------------------------------------------
// gemm: D = alpha * op(A) * op(B) + beta * C, CV_32F / CV_64F, following modules/core/src/matmul.dispatch.cpp
enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };
inline Mat gemm_eval(const Mat& A, const Mat& B, double alpha, const Mat* C, double beta, int flags) {
  const bool ta = flags & GEMM_1_T, tb = flags & GEMM_2_T;
  const int ar = ta ? A.cols : A.rows, len = ta ? A.rows : A.cols, bc = tb ? B.rows : B.cols;
  assert((tb ? B.cols : B.rows) == len && A.type() == B.type());
  Mat D(ar, bc, A.type());
  auto a = [&](int i, int k) { return ta ? A.get_d(k, i) : A.get_d(i, k); };
  auto b = [&](int k, int j) { return tb ? B.get_d(j, k) : B.get_d(k, j); };
  auto c = [&](int i, int j) { return C ? ((flags & GEMM_3_T) ? C->get_d(j, i) : C->get_d(i, j)) : 0.0; };
  const bool small = flags == 0 && 2 <= len && len <= 4 && (len == bc || len == ar);
  for (int i = 0; i < ar; ++i)
    for (int j = 0; j < bc; ++j) {
      if (A.depth() == CV_32F && small) {
        float t = (float)a(i, 0) * (float)b(0, j);   // float products, float sum, left to right
        for (int k = 1; k < len; ++k) t = t + (float)a(i, k) * (float)b(k, j);
        D.at<float>(i, j) = C ? (float)((double)t * alpha + c(i, j) * beta) : (float)((double)t * alpha);
      } else {
        double s = 0;
        for (int k = 0; k < len; ++k) s += a(i, k) * b(k, j);
        D.set_d(i, j, C ? s * alpha + c(i, j) * beta : s * alpha);
      }
    }
  return D;
}

struct Mat::TExpr {   // alpha * A^T, unevaluated
  Mat a; double alpha;
  operator Mat() const {   // transpose, then convertTo(alpha) when scaled (float product for CV_32F)
    Mat d(a.cols, a.rows, a.type());
    for (int i = 0; i < a.rows; ++i)
      for (int j = 0; j < a.cols; ++j) {
        if (alpha == 1.0) d.set_d(j, i, a.get_d(i, j));
        else if (a.depth() == CV_32F) d.at<float>(j, i) = a.at<float>(i, j) * (float)alpha;
        else d.set_d(j, i, a.get_d(i, j) * alpha);
      }
    return d;
  }
  TExpr operator-() const { return TExpr{a, -alpha}; }
  Mat rowRange(int s, int e) const { return Mat(*this).rowRange(s, e); }
};
inline Mat::TExpr Mat::t() const { return TExpr{*this, 1.0}; }

struct GemmExpr {   // alpha * op(A) * op(B), unevaluated so that "+ C" folds into the same gemm call
  Mat a, b; double alpha; int flags;
  operator Mat() const { return gemm_eval(a, b, alpha, nullptr, 0, flags); }
  GemmExpr operator-() const { return GemmExpr{a, b, -alpha, flags}; }
  template <typename T> T at(int i, int j = 0) const { return Mat(*this).at<T>(i, j); }
};
inline GemmExpr operator*(const Mat& a, const Mat& b) { return GemmExpr{a, b, 1.0, 0}; }
inline GemmExpr operator*(const Mat::TExpr& a, const Mat& b) { return GemmExpr{a.a, b, a.alpha, GEMM_1_T}; }
inline GemmExpr operator*(const Mat& a, const Mat::TExpr& b) { return GemmExpr{a, b.a, b.alpha, GEMM_2_T}; }
inline GemmExpr operator*(const GemmExpr& a, const Mat& b) { return GemmExpr{Mat(a), b, 1.0, 0}; }
inline Mat operator+(const GemmExpr& g, const Mat& c) { return gemm_eval(g.a, g.b, g.alpha, &c, 1.0, g.flags); }
inline Mat operator+(const Mat& c, const GemmExpr& g) { return gemm_eval(g.a, g.b, g.alpha, &c, 1.0, g.flags); }
inline Mat operator-(const GemmExpr& g, const Mat& c) { return gemm_eval(g.a, g.b, g.alpha, &c, -1.0, g.flags); }

// element-wise (cv::add / subtract / scaleAdd on CV_32F: float arithmetic; on CV_64F: double)
template <typename F> inline Mat elementwise(const Mat& a, const Mat& b, F f) {
  assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
  Mat d(a.rows, a.cols, a.type());
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols * a.channels(); ++j) {
      if (a.depth() == CV_32F) d.ptr<float>(i)[j] = f(a.ptr<float>(i)[j], b.ptr<float>(i)[j]);
      else d.set_d(i, j, f(a.get_d(i, j), b.get_d(i, j)));
    }
  return d;
}
inline Mat operator+(const Mat& a, const Mat& b) { return elementwise(a, b, [](auto x, auto y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return elementwise(a, b, [](auto x, auto y) { return x - y; }); }
inline Mat scaled(const Mat& a, double s) {   // MatExpr alpha*A -> A.convertTo(type, alpha): for CV_32F the scale is
  Mat d(a.rows, a.cols, a.type());            // narrowed to float and the product formed in float (cvtScale 32f->32f)
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols * a.channels(); ++j) {
      if (a.depth() == CV_32F) d.ptr<float>(i)[j] = a.ptr<float>(i)[j] * (float)s;
      else d.set_d(i, j, a.get_d(i, j) * s);
    }
  return d;
}
inline Mat operator*(const Mat& a, double s) { return scaled(a, s); }
inline Mat operator*(double s, const Mat& a) { return scaled(a, s); }
inline Mat operator*(double s, const MatInit& a) { return scaled(Mat(a), s); }
inline Mat operator/(const Mat& a, double s) { return scaled(a, 1.0 / s); }   // MatExpr: A/s == A*(1/s)
inline Mat::TExpr operator*(double s, const Mat::TExpr& t) { return Mat::TExpr{t.a, t.alpha * s}; }
inline Mat::TExpr operator*(const Mat::TExpr& t, double s) { return Mat::TExpr{t.a, t.alpha * s}; }
inline Mat operator-(const Mat& a) { return scaled(a, -1.0); }

inline double norm(const Mat& a, int normType = NORM_L2) {
  double s = 0;
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols * a.channels(); ++j) {
      double v = a.get_d(i, j);
      if (normType == NORM_L2) s += v * v;
      else if (normType == NORM_L1) s += std::fabs(v);
      else s = std::max(s, std::fabs(v));
    }
  return normType == NORM_L2 ? std::sqrt(s) : s;
}
inline double norm(const Mat& a, const Mat& b, int normType = NORM_L2) { return norm(a - b, normType); }
inline Scalar sum(const Mat& a) {   // cv::sum: per-channel sums (integer inputs: exact in double whatever OpenCV's blocking)
  Scalar s;
  const int cn = a.channels();
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols * cn; ++j) s.val[j % cn] += a.get_d(i, j);
  return s;
}
inline double norm(const GemmExpr& g) { return norm(Mat(g)); }

inline Mat Mat::inv() const {   // only used by out-of-path code; Gauss-Jordan in double
  assert(rows == cols);
  const int n = rows;
  std::vector<double> m((size_t)n * 2 * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) m[(size_t)i * 2 * n + j] = get_d(i, j);
    m[(size_t)i * 2 * n + n + i] = 1;
  }
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r) if (std::fabs(m[(size_t)r * 2 * n + c]) > std::fabs(m[(size_t)p * 2 * n + c])) p = r;
    for (int j = 0; j < 2 * n; ++j) std::swap(m[(size_t)c * 2 * n + j], m[(size_t)p * 2 * n + j]);
    double d = m[(size_t)c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) m[(size_t)c * 2 * n + j] /= d;
    for (int r = 0; r < n; ++r) if (r != c) {
      double f = m[(size_t)r * 2 * n + c];
      for (int j = 0; j < 2 * n; ++j) m[(size_t)r * 2 * n + j] -= f * m[(size_t)c * 2 * n + j];
    }
  }
  Mat d(n, n, type());
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) d.set_d(i, j, m[(size_t)i * 2 * n + n + j]);
  return d;
}

// ---------------------------------------------------------------------------------------------------------------------
// InputArray / OutputArray: thin proxies over Mat (enough for ORBextractor::operator())
class _InputArray {
 public:
  Mat* m; Mat own;
  _InputArray() : m(nullptr) {}
  _InputArray(const Mat& mm) : m(const_cast<Mat*>(&mm)) {}
  _InputArray(const MatInit& e) : m(&own), own(e) {}
  bool empty() const { return !m || m->empty(); }
  Mat getMat() const { return m ? *m : Mat(); }
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& mm) : _InputArray(mm) {}
  void create(int r, int c, int type) const { m->create(r, c, type); }
  void release() const { if (m) m->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline const _OutputArray& noArray() { static _OutputArray a; return a; }

// ---------------------------------------------------------------------------------------------------------------------
// image primitives (cv_prims.hpp holds the arithmetic; these adapt Mat views to it)
inline void resize(InputArray src_, OutputArray dst_, Size dsize, double = 0, double = 0, int interp = INTER_LINEAR) {
  Mat src = src_.getMat();
  assert(interp == INTER_LINEAR && src.type() == CV_8UC1);
  Mat dst = dst_.getMat();
  if (dst.rows != dsize.height || dst.cols != dsize.width || dst.type() != CV_8UC1) {
    dst_.create(dsize.height, dsize.width, CV_8UC1);
    dst = dst_.getMat();
  }
  cvprim::resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}
inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType) {
  Mat src = src_.getMat();
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && src.type() == CV_8UC1);
  Mat dst = dst_.getMat();
  // OpenCV: dst.create(rows+top+bottom, cols+left+right) is a no-op when dst already has that shape (the reference
  // relies on it: `temp` stays the parent buffer of the pyramid-level ROI, src/ORBextractor.cc:1128-1143)
  if (dst.rows != src.rows + top + bottom || dst.cols != src.cols + left + right || dst.type() != src.type()) {
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    dst = dst_.getMat();
  }
  cvprim::copy_make_border_101_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step, top, bottom, left, right);
}
inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sx, double sy, int borderType) {
  Mat src = src_.getMat();
  assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
  assert(src.type() == CV_8UC1);
  Mat tmp(src.rows, src.cols, CV_8UC1);
  cvprim::gaussian7_u8(src.data, src.cols, src.rows, (int)src.step, tmp.data, (int)tmp.step);
  Mat dst = dst_.getMat();
  if (dst.rows != src.rows || dst.cols != src.cols) { dst_.create(src.rows, src.cols, CV_8UC1); dst = dst_.getMat(); }
  for (int i = 0; i < src.rows; ++i) memcpy(dst.ptr(i), tmp.ptr(i), (size_t)src.cols);
}
inline void FAST(InputArray img_, std::vector<KeyPoint>& kps, int threshold, bool nonmax = true) {
  Mat img = img_.getMat();
  assert(img.type() == CV_8UC1);
  std::vector<cvprim::FastKp> out;
  cvprim::fast9_roi(img.data, img.cols, img.rows, (int)img.step, threshold, out, nonmax);
  kps.clear();
  for (const auto& k : out) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1.f, (float)k.score));
}
inline float fastAtan2(float y, float x) { return cvprim::fast_atan2_deg(y, x); }

struct KeyPointsFilter {   // only referenced by ComputeKeyPointsOld (dead code in the reference, src/ORBextractor.cc:866)
  static void retainBest(std::vector<KeyPoint>& kps, int n) {
    if (n >= 0 && (int)kps.size() > n) {
      std::stable_sort(kps.begin(), kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
      kps.resize(n);
    }
  }
};

// cv::undistortPoints(src Nx1 CV_32FC2, dst, K, dist(k1 k2 p1 p2 [k3]), R = empty, P = K): the classic fixed-point
// iteration (5 iterations, no termination test), computed in double, result stored as float.
inline void undistortPoints(InputArray src_, OutputArray dst_, InputArray K_, InputArray dist_, InputArray /*R*/, InputArray P_) {
  Mat src = src_.getMat(), K = K_.getMat(), D = dist_.getMat(), P = P_.getMat();
  const double fx = K.get_d(0, 0), fy = K.get_d(1, 1), cx = K.get_d(0, 2), cy = K.get_d(1, 2);
  const double ifx = 1. / fx, ify = 1. / fy;
  double k[5] = {0, 0, 0, 0, 0};
  const int nk = (int)D.total();
  for (int i = 0; i < std::min(nk, 5); ++i) k[i] = D.rows == 1 ? D.get_d(0, i) : D.get_d(i, 0);
  const double pfx = P.get_d(0, 0), pfy = P.get_d(1, 1), pcx = P.get_d(0, 2), pcy = P.get_d(1, 2);
  Mat out(src.rows, src.cols, src.type());
  (void)ifx; (void)ify; (void)pfx; (void)pfy; (void)pcx; (void)pcy;   // P == K at the reference's call sites
  for (int i = 0; i < src.rows; ++i) {
    double x, y;
    cvprim::undistort_point(src.ptr<float>(i)[0], src.ptr<float>(i)[1], fx, fy, cx, cy, k, &x, &y);
    out.ptr<float>(i)[0] = (float)x;
    out.ptr<float>(i)[1] = (float)y;
  }
  Mat dst = dst_.getMat();
  if (dst.data == src.data || (dst.rows == out.rows && dst.cols == out.cols && dst.type() == out.type()))
    out.copyTo(dst);
  else { dst_.create(out.rows, out.cols, out.type()); Mat d2 = dst_.getMat(); out.copyTo(d2); }
}

}  // namespace cv
