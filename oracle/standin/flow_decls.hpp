// Declarations (no definitions) of the three OpenCV calls shim/Flow.h leaves on the host, so that the header can be
// syntax-checked against the stand-in cv:: types where OpenCV's C++ headers are absent (tests/test_abi_cpu.py).
#pragma once
#include <opencv2/opencv.hpp>
namespace cv {
void pyrDown(const Mat& src, Mat& dst, const Size& dstsize);
void calcOpticalFlowFarneback(const Mat& prev, const Mat& next, Mat& flow, double pyr_scale, int levels, int winsize,
                              int iterations, int poly_n, double poly_sigma, int flags);
void warpPerspective(const Mat& src, Mat& dst, const Mat& M, Size dsize);
}  // namespace cv
