// oracle/refperfect_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// C entry point around the `perfect` tree's OWN translation units, compiled unmodified from /root/reference/perfect into
// oracle/_ref/librefperfect.so (oracle/Makefile, target `perfect`): perfect/src/Frame.cc, KeyFrame.cc, MapPoint.cc, Map.cc,
// ORBextractor.cc, ORBmatcher.cc.  It builds the two RGB-D Frames the perfect variant can build from one image pair -- the
// plain constructor (perfect/src/Frame.cc:255-322) and the MASKED one (:328-427), whose keypoint loop (:356-377) is what
// dynm_filter_keypoints replaces -- and hands back the (distorted) keypoints and descriptors each of them keeps.
// tests/test_dynmask_cpu.py asserts oracle/dynmask_py.filter_keypoints(mask, plain) == masked, bit for bit: that pins the
// oracle of that loop to reference code.  Nothing here restates reference logic.
//
// The quad-tree's list nodes come from a bump arena (same device as refsrc_harness.cpp: DistributeOctTree orders equal-size
// nodes by heap address, perfect/src/ORBextractor.cc, so two extractions of one image only agree in ORDER when "higher
// address" means "created later" in both).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <list>
#include <mutex>
#include <new>
#include <vector>

#define private public
#define protected public
#include "ORBextractor.h"
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "Map.h"
#include "KeyFrameDatabase.h"
#undef private
#undef protected

namespace ORB_SLAM2 {
// out-of-path symbols the six translation units leave undefined: the key-frame database (KeyFrame::SetBadFlag) and the two
// quaternion helpers of the perfect tree's map save / load (perfect/src/Map.cc:163,354; Eigen is absent).  Never reached.
void KeyFrameDatabase::erase(KeyFrame*) {}
std::vector<float> Converter::toQuaternion(const cv::Mat&) { abort(); }
void Converter::RmatOfQuat(cv::Mat&, const cv::Mat&) { abort(); }
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

namespace {
const size_t kNodeBytes = sizeof(std::_List_node<ExtractorNode>);
const size_t kArenaBytes = 16u << 20;
char* g_arena = nullptr;
size_t g_used = 0;
bool g_on = false;
std::mutex g_mu;
void arena_reset() {
  if (!g_arena) g_arena = (char*)malloc(kArenaBytes);
  g_used = 0;
  g_on = g_arena != nullptr;
}
}  // namespace
void* operator new(size_t n) {
  if (g_on && n == kNodeBytes && g_used + ((n + 15) & ~size_t(15)) <= kArenaBytes) {
    void* p = g_arena + g_used;
    g_used += (n + 15) & ~size_t(15);
    return p;
  }
  void* p = malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void operator delete(void* p) noexcept {
  if (g_arena && (char*)p >= g_arena && (char*)p < g_arena + kArenaBytes) return;
  free(p);
}
void operator delete(void* p, size_t) noexcept { operator delete(p); }

extern "C" {

// gray u8, depth f32 metres, mask u8, all rows x cols.  kps_* : cap x 28 bytes (cv::KeyPoint), desc_* : cap x 32.
// -> 0, or -2 when cap is too small.
int refperfect_frames(const uint8_t* gray, const float* depth, const uint8_t* mask, int rows, int cols, int nfeatures,
                      float scaleFactor, int nlevels, int iniTh, int minTh, float fx, float fy, float cx, float cy, float bf,
                      void* kps_plain, uint8_t* desc_plain, int* n_plain, void* kps_masked, uint8_t* desc_masked,
                      int* n_masked, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
  K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy;
  cv::Mat D = cv::Mat::zeros(4, 1, CV_32F);
  cv::Mat g(rows, cols, CV_8UC1, (void*)gray), d(rows, cols, CV_32F, (void*)depth), m(rows, cols, CV_8UC1, (void*)mask);
  const float thDepth = bf * 40.0f / fx;
  Frame::mbInitialComputations = true;
  arena_reset();
  Frame A(g, d, 0.0, &ex, nullptr, K, D, bf, thDepth);          // perfect/src/Frame.cc:255
  arena_reset();
  Frame B(g, d, m, 0.0, &ex, nullptr, K, D, bf, thDepth);       // perfect/src/Frame.cc:328 (masked)
  g_on = false;
  *n_plain = A.N;
  *n_masked = B.N;
  if (A.N > cap || B.N > cap) return -2;
  for (int i = 0; i < A.N; ++i) {
    memcpy((char*)kps_plain + 28 * (size_t)i, &A.mvKeys[i], 28);
    memcpy(desc_plain + 32 * (size_t)i, A.mDescriptors.ptr(i), 32);
  }
  for (int i = 0; i < B.N; ++i) {
    memcpy((char*)kps_masked + 28 * (size_t)i, &B.mvKeys[i], 28);
    memcpy(desc_masked + 32 * (size_t)i, B.mDescriptors.ptr(i), 32);
  }
  return 0;
}

}  // extern "C"
