// oracle/standin/Converter.h (TEST INFRASTRUCTURE): shadows the reference's include/Converter.h for the .cc files that
// include it with quotes (src/Frame.cc:47, src/KeyFrame.cc:30).  The real header drags in Eigen + g2o (absent); the only
// member those files use is toDescriptorVector (src/Converter.cc:13-22: one row Mat per descriptor), restated here.
#ifndef CONVERTER_H
#define CONVERTER_H
#include <opencv2/core/core.hpp>
#include <vector>
namespace ORB_SLAM2 {
class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors) {
    std::vector<cv::Mat> vDesc;
    vDesc.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) vDesc.push_back(Descriptors.row(j));
    return vDesc;
  }
};
}  // namespace ORB_SLAM2
#endif
