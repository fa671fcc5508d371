// Compiles the shim headers against the stand-in cv:: types and exercises them when a GPU is present.
// Built and run by tests/test_abi_cpu.py (compile + link only on the CPU box) and tests/test_shim_gpu.py.
#define B200_SHIM_STANDIN
#include <cstdio>
#include <map>

#include "ORBextractor.h"
#include "pointcloudmapping.h"
#include "ORBmatcher.h"

// minimal stand-ins with the member names the reference's Frame / MapPoint expose (include/Frame.h, MapPoint.h)
struct MockMapPoint {
  cv::Mat pos{3, 1, CV_32F}, desc{1, 32, CV_8UC1};
  int obs = 1;
  bool mbTrackInView = true;
  float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 1;
  int mnTrackScaleLevel = 0;
  cv::Mat GetWorldPos() { return pos; }
  cv::Mat GetDescriptor() { return desc; }
  int Observations() { return obs; }
  bool isBad() { return false; }
};
typedef std::map<unsigned int, std::vector<unsigned int> > MockFeatureVector;   // DBoW2::FeatureVector's shape
struct MockFrame {
  int N = 0;
  MockFeatureVector mFeatVec;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  std::vector<float> mvuRight, mvScaleFactors;
  std::vector<MockMapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  cv::Mat mDescriptors, mTcw{4, 4, CV_32F};
  float fx = 535.4f, fy = 539.2f, cx = 320.1f, cy = 247.6f, mbf = 40.f, mb = 40.f / 535.4f;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
float MockFrame::mnMinX = 0, MockFrame::mnMaxX = 640, MockFrame::mnMinY = 0, MockFrame::mnMaxY = 480;
struct MockKeyFrame {   // the members SearchByBoW reads (include/KeyFrame.h)
  std::vector<cv::KeyPoint> mvKeysUn;
  cv::Mat mDescriptors;
  MockFeatureVector mFeatVec;
  std::vector<MockMapPoint*> mps;
  std::vector<MockMapPoint*> GetMapPointMatches() { return mps; }
};
typedef ORB_SLAM2::ORBmatcherT<MockFrame, MockMapPoint> MockMatcher;

int main(int argc, char** argv) {
  if (argc > 1000) {   // never taken: instantiates the BoW overloads of the matcher shim so that they are compile-checked
    MockMatcher m(0.7f, true);
    MockKeyFrame k1, k2;
    MockFrame f;
    std::vector<MockMapPoint*> out;
    return m.SearchByBoW(&k1, f, out) + m.SearchByBoW(&k1, &k2, out);
  }
  if (b200orb_device_count() == 0) {
    try {
      ORB_SLAM2::ORBextractor e(1000, 1.2f, 8, 20, 7);
      std::printf("UNEXPECTED: constructed without a GPU\n");
      return 2;
    } catch (const std::exception& ex) {
      std::printf("no GPU -> loud failure: %s\n", ex.what());
      return 0;
    }
  }
  ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
  cv::Mat img(480, 640, CV_8UC1);
  unsigned s = 12345u;
  for (int i = 0; i < 480 * 640; ++i) { s = s * 1664525u + 1013904223u; img.data[i] = (uint8_t)((s >> 24) & 0xff); }
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc, mask;
  ex(img, mask, kps, desc);
  std::printf("keypoints %zu desc %dx%d levels %d sf1 %.4f\n", kps.size(), desc.rows, desc.cols, ex.GetLevels(),
              ex.GetScaleFactors()[1]);
  ex.SyncImagePyramid();
  std::printf("pyramid L7 %dx%d\n", ex.mvImagePyramid[7].cols, ex.mvImagePyramid[7].rows);
  {   // matcher shim: a frame matched against itself through MapPoints placed on the z = 2 plane
    MockFrame F, L;
    std::vector<MockMapPoint> mps(kps.size());
    for (MockFrame* f : {&F, &L}) {
      f->N = (int)kps.size(); f->mvKeys = kps; f->mvKeysUn = kps; f->mvuRight.assign(kps.size(), -1.f);
      f->mvScaleFactors = ex.GetScaleFactors(); f->mvpMapPoints.assign(kps.size(), nullptr); f->mvbOutlier.assign(kps.size(), false);
      f->mDescriptors = desc;
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) f->mTcw.at<float>(r, c) = (r == c) ? 1.f : 0.f;
    }
    for (size_t i = 0; i < kps.size(); ++i) {
      mps[i].pos.at<float>(0, 0) = (kps[i].pt.x - F.cx) * 2.f / F.fx;
      mps[i].pos.at<float>(1, 0) = (kps[i].pt.y - F.cy) * 2.f / F.fy;
      mps[i].pos.at<float>(2, 0) = 2.f;
      std::memcpy(mps[i].desc.ptr(0), desc.ptr((int)i), 32);
      L.mvpMapPoints[i] = &mps[i];
    }
    MockMatcher m(0.9f, true);
    const int nm = m.SearchByProjection(F, L, 15.f, false);
    int self = 0;
    for (size_t i = 0; i < kps.size(); ++i) self += (F.mvpMapPoints[i] == &mps[i]);
    std::printf("matcher shim: %d matches, %d keypoints matched to themselves\n", nm, self);
    if (self < (int)kps.size() * 9 / 10) return 3;
  }
  PointCloudMapping pcm(0.05);
  cv::Mat depth(480, 640, CV_32F), rgb(480, 640 * 3, CV_8UC1);
  for (int r = 0; r < 480; ++r) for (int c = 0; c < 640; ++c) depth.at<float>(r, c) = 1.5f + 0.001f * c;
  float Tcw[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  pcm.insertKeyFrame(Tcw, 535.4f, 539.2f, 320.1f, 247.6f, img, depth, rgb);
  pcm.shutdown();
  std::printf("leaves %lld\n", (long long)pcm.numLeaves());
  PointCloudMappingT pcmT(0.04);                       // T variant: accumulated cloud + whole-map VoxelGrid refilter
  pcmT.insertKeyFrame(Tcw, 535.4f, 539.2f, 320.1f, 247.6f, img, depth, rgb);
  const long long raw = pcmT.globalMapSize();
  pcmT.update();
  pcmT.shutdown();
  std::printf("global map %lld -> %lld points\n", raw, pcmT.globalMapSize());
  return (kps.size() >= 1000 && desc.rows == (int)kps.size() && pcm.numLeaves() > 0 && raw == 480 * 640 &&
          pcmT.globalMapSize() > 0 && pcmT.globalMapSize() < raw) ? 0 : 1;
}
