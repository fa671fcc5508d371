// oracle/standin/cv_prims.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Integer / float models of the OpenCV primitives the reference's extractor delegates to:
// cv::resize(INTER_LINEAR, 8UC1), copyMakeBorder(REFLECT_101), GaussianBlur(7x7, sigma 2), FAST(TYPE_9_16, nonmax),
// fastAtan2.  OpenCV is a third-party dependency that is absent from /root/reference and from this image's C++
// toolchain; each model is compared bit for bit with the real cv2 4.13.0 wheel in tests/test_oracle_cpu.py.
// ONE copy, two users: oracle/orb_ref.cpp (the restatement) and oracle/standin/cv_standin.hpp (the cv:: stand-in the
// reference's own sources are compiled against for oracle/_ref/).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace cvprim {

inline int cvRoundf(float v) { return (int)lrintf(v); }   // A.1: round-half-even (default FE mode)
inline int cvRoundd(double v) { return (int)lrint(v); }

// ---------------------------------------------------------------------------------------------
// A.2  cv::resize(INTER_LINEAR) for 8UC1 (call site src/ORBextractor.cc:1134)
// ---------------------------------------------------------------------------------------------
inline void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                      int dstride) {
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> xa(2 * dw), ya(2 * dh);
  auto coeffs = [](int dn, int sn, int* ofs, short* ab) {
    double scale = 1.0 / ((double)dn / sn);
    for (int d = 0; d < dn; ++d) {
      float f = (float)((d + 0.5) * scale - 0.5);
      int i = (int)std::floor(f);
      f -= (float)i;
      if (i < 0) { i = 0; f = 0.f; }
      if (i >= sn - 1) { i = sn - 1; f = 0.f; }
      ofs[d] = i;
      ab[2 * d] = (short)cvRoundf((1.f - f) * 2048.f);
      ab[2 * d + 1] = (short)cvRoundf(f * 2048.f);
    }
  };
  coeffs(dw, sw, xofs.data(), xa.data());
  coeffs(dh, sh, yofs.data(), ya.data());
  std::vector<int> r0(dw), r1(dw);
  for (int dy = 0; dy < dh; ++dy) {
    int sy0 = yofs[dy], sy1 = std::min(sy0 + 1, sh - 1);
    const uint8_t* s0 = src + (size_t)sy0 * sstride;
    const uint8_t* s1 = src + (size_t)sy1 * sstride;
    for (int dx = 0; dx < dw; ++dx) {
      int x0 = xofs[dx], x1 = std::min(x0 + 1, sw - 1);
      r0[dx] = s0[x0] * xa[2 * dx] + s0[x1] * xa[2 * dx + 1];
      r1[dx] = s1[x0] * xa[2 * dx] + s1[x1] * xa[2 * dx + 1];
    }
    int b0 = ya[2 * dy], b1 = ya[2 * dy + 1];
    uint8_t* o = dst + (size_t)dy * dstride;
    for (int dx = 0; dx < dw; ++dx)
      o[dx] = (uint8_t)((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
  }
}

inline int reflect101(int p, int n) {  // BORDER_REFLECT_101
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

// ---------------------------------------------------------------------------------------------
// A.3  cv::GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101 (call site src/ORBextractor.cc:1095)
// ---------------------------------------------------------------------------------------------
inline void gaussian7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  static const int q[7] = {18, 34, 48, 56, 48, 34, 18};
  std::vector<uint16_t> hb((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = src + (size_t)y * sstride;
    uint16_t* o = &hb[(size_t)y * w];
    for (int x = 0; x < w; ++x) {
      if (x >= 3 && x < w - 3) {
        o[x] = (uint16_t)(18 * (s[x - 3] + s[x + 3]) + 34 * (s[x - 2] + s[x + 2]) + 48 * (s[x - 1] + s[x + 1]) + 56 * s[x]);
      } else {
        int acc = 0;
        for (int k = 0; k < 7; ++k) acc += q[k] * s[reflect101(x + k - 3, w)];
        o[x] = (uint16_t)acc;
      }
    }
  }
  for (int y = 0; y < h; ++y) {
    uint8_t* o = dst + (size_t)y * dstride;
    const uint16_t* r[7];
    for (int k = 0; k < 7; ++k) r[k] = &hb[(size_t)reflect101(y + k - 3, h) * w];
    for (int x = 0; x < w; ++x) {
      const uint32_t acc = 18u * (r[0][x] + r[6][x]) + 34u * (r[1][x] + r[5][x]) + 48u * (r[2][x] + r[4][x]) + 56u * r[3][x];
      o[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// A.4  cv::FAST(TYPE_9_16, nonmax=true) on an ROI (call sites src/ORBextractor.cc:818,823)
// ---------------------------------------------------------------------------------------------
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// m(p) = max over the 16 arcs of 9 contiguous ring pixels of max(min d, min -d); corner iff m > t.
inline int fast_m(const uint8_t* p, int stride) {
  int d[25];
  int c = p[0];
  for (int k = 0; k < 16; ++k) d[k] = c - p[kRingDy[k] * stride + kRingDx[k]];
  for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
  int best = 0;
  for (int s = 0; s < 16; ++s) {
    int mn = d[s], mx = d[s];
    for (int j = 1; j < 9; ++j) {
      mn = std::min(mn, d[s + j]);
      mx = std::max(mx, d[s + j]);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

struct FastKp { int x, y, score; };

// corner test at threshold t without the score: 16-bit masks of ring pixels brighter than c+t / darker than
// c-t and a 9-contiguous-bits test on the doubled mask (equivalent to m > t; the exact m is only evaluated for
// corners, like OpenCV evaluates cornerScore only for detected corners).
inline bool fast_is_corner(const uint8_t* p, int stride, int t) {
  const int c = p[0], hi = c + t, lo = c - t;
  const int r0 = p[3 * stride], r8 = p[-3 * stride];
  if (!((r0 > hi) | (r8 > hi) | (r0 < lo) | (r8 < lo))) return false;   // any 9-arc holds pixel 0 or pixel 8
  const int r4 = p[3], r12 = p[-3];
  if (!((r4 > hi) | (r12 > hi) | (r4 < lo) | (r12 < lo))) return false;
  uint32_t mb = 0, md = 0;
  for (int k = 0; k < 16; ++k) {
    const int v = p[kRingDy[k] * stride + kRingDx[k]];
    mb |= (uint32_t)(v > hi) << k;
    md |= (uint32_t)(v < lo) << k;
  }
  auto run9 = [](uint32_t m) {
    m |= m << 16;
    m &= m >> 1;   // runs >= 2
    m &= m >> 2;   // runs >= 4
    m &= m >> 4;   // runs >= 8
    m &= m >> 1;   // runs >= 9
    return m != 0;
  };
  return run9(mb) || run9(md);
}

inline void fast9_roi(const uint8_t* roi, int rw, int rh, int stride, int t, std::vector<FastKp>& out,
               std::vector<int>& sc /*scratch rw*rh*/) {
  out.clear();
  if (rw < 7 || rh < 7) return;
  sc.assign((size_t)rw * rh, 0);   // the 3-px ROI margin never scores (stays 0)
  bool any = false;
  for (int y = 3; y < rh - 3; ++y)
    for (int x = 3; x < rw - 3; ++x) {
      const uint8_t* p = roi + (size_t)y * stride + x;
      if (!fast_is_corner(p, stride, t)) continue;
      sc[(size_t)y * rw + x] = fast_m(p, stride) - 1;   // m > t >= 0
      any = true;
    }
  if (!any) return;
  for (int y = 3; y < rh - 3; ++y)
    for (int x = 3; x < rw - 3; ++x) {
      const int* r = &sc[(size_t)y * rw + x];
      const int s = r[0];   // strict 3x3 maximum; non-corners hold 0, so s > 0 follows
      if (s > r[-1] && s > r[1] && s > r[-rw - 1] && s > r[-rw] && s > r[-rw + 1] && s > r[rw - 1] &&
          s > r[rw] && s > r[rw + 1])
        out.push_back({x, y, s});
    }
}

// ---------------------------------------------------------------------------------------------
// A.5  cv::fastAtan2 (call site src/ORBextractor.cc:87)
// ---------------------------------------------------------------------------------------------
inline float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + 2.2204460492503131e-16f);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + 2.2204460492503131e-16f);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// cv::FAST as called by the reference: own scratch, optional NMS (nonmax=false returns every corner, row-major)
inline void fast9_roi(const uint8_t* roi, int rw, int rh, int stride, int t, std::vector<FastKp>& out, bool nonmax) {
  std::vector<int> sc;
  if (nonmax) { fast9_roi(roi, rw, rh, stride, t, out, sc); return; }
  out.clear();
  if (rw < 7 || rh < 7) return;
  for (int y = 3; y < rh - 3; ++y)
    for (int x = 3; x < rw - 3; ++x) {
      const uint8_t* p = roi + (size_t)y * stride + x;
      if (fast_is_corner(p, stride, t)) out.push_back({x, y, fast_m(p, stride) - 1});
    }
}

// cv::copyMakeBorder(BORDER_REFLECT_101 [| BORDER_ISOLATED]) for 8UC1; src may live inside dst (the reference's
// pyramid level is an ROI of the bordered buffer it is copied into, src/ORBextractor.cc:1128-1143): the interior is
// moved first, then the left/right borders of the interior rows, then whole rows for top/bottom.
inline void copy_make_border_101_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int top,
                                    int bottom, int left, int right) {
  for (int y = 0; y < h; ++y) {
    uint8_t* o = dst + (size_t)(y + top) * dstride;
    if (o + left != src + (size_t)y * sstride) memmove(o + left, src + (size_t)y * sstride, (size_t)w);
  }
  for (int y = 0; y < h; ++y) {
    uint8_t* o = dst + (size_t)(y + top) * dstride;
    for (int x = 0; x < left; ++x) o[x] = o[left + reflect101(x - left, w)];
    for (int x = 0; x < right; ++x) o[left + w + x] = o[left + reflect101(w + x, w)];
  }
  const int W = w + left + right;
  for (int y = 0; y < top; ++y) memcpy(dst + (size_t)y * dstride, dst + (size_t)(top + reflect101(y - top, h)) * dstride, (size_t)W);
  for (int y = 0; y < bottom; ++y)
    memcpy(dst + (size_t)(top + h + y) * dstride, dst + (size_t)(top + reflect101(h + y, h)) * dstride, (size_t)W);
}

// cv::undistortPoints(src, dst, K, dist(k1 k2 p1 p2 k3), R = I, P = K) for one point: the classic fixed-point iteration, 5
// iterations (OpenCV's default termination criteria), all in double; the caller stores the result as float.
inline void undistort_point(double u, double v, double fx, double fy, double cx, double cy, const double k[5], double* xo,
                            double* yo) {
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = 1. / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) { x = x0; y = y0; break; }   // OpenCV gives up and keeps the normalised input
    const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
    const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  *xo = x * fx + cx;
  *yo = y * fy + cy;
}

}  // namespace cvprim
