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
  // the perfect tree's map save / load (perfect/src/Map.cc:163,354) names two quaternion helpers; declared only -- Eigen is
  // absent and the path never calls them (oracle/refperfect_harness.cpp defines aborting bodies for the link)
  static std::vector<float> toQuaternion(const cv::Mat& M);
  void RmatOfQuat(cv::Mat& M, const cv::Mat& q);
};
}  // namespace ORB_SLAM2
#endif
