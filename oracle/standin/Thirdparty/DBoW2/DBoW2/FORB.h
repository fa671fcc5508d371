// oracle/standin: DBoW2::FORB (TEST INFRASTRUCTURE): 32-byte ORB descriptor functions (Hamming distance).
#pragma once
#include <opencv2/core/core.hpp>
#include <cstdint>
namespace DBoW2 {
class FORB {
 public:
  typedef cv::Mat TDescriptor;
  typedef const TDescriptor* pDescriptor;
  static const int L = 32;
  static int distance(const TDescriptor& a, const TDescriptor& b) {   // 256-bit Hamming distance
    const uint32_t* pa = a.ptr<uint32_t>();
    const uint32_t* pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; ++i) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
  }
};
}  // namespace DBoW2
