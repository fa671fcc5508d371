// shim/pointcloudmapping.h -- PointCloudMapping with the reference's public surface (include/pointcloudmapping.h:50-56,
// perfect/include/pointcloudmapping.h:23-28) in its occupancy mode: insertKeyFrame() forwards the keyframe's pose,
// intrinsics, depth and colour image to ocm_insert_keyframe (MapDrawer::GeneratePointCloud + InsertScan semantics,
// perfect/src/MapDrawer.cc:641-675,946-1025).  The reference's insertKeyFrame only clones + notifies and a worker
// thread does the work (src/pointcloudmapping.cc:116-128,362-525); here the GPU stream plays the worker: the call
// enqueues and returns, shutdown() drains it.
#ifndef POINTCLOUDMAPPING_H
#define POINTCLOUDMAPPING_H

#include <stdexcept>
#include <string>
#include <vector>

#include <cstdlib>
#include <opencv2/opencv.hpp>

#include "b200orb.h"

class PointCloudMapping {
 public:
  explicit PointCloudMapping(double resolution_) : resolution(resolution_) {
    OcmParams p;
    ocm_default_params(&p);
    p.resolution = resolution_;
    int dev = 0;
    if (const char* e = std::getenv("B200ORB_DEVICE")) dev = std::atoi(e);
    if (ocm_create(&p, dev, &h_) != B200ORB_OK)
      throw std::runtime_error(std::string("PointCloudMapping(B200): ") + b200orb_last_error());
  }
  ~PointCloudMapping() { ocm_destroy(h_); }
  PointCloudMapping(const PointCloudMapping&) = delete;
  PointCloudMapping& operator=(const PointCloudMapping&) = delete;

  // Reference: insertKeyFrame(KeyFrame* kf, cv::Mat& color, cv::Mat& depth, cv::Mat& imgRGB) -- the KeyFrame supplies
  // GetPose() (Tcw, CV_32F 4x4) and fx/fy/cx/cy.  Template keeps this header free of KeyFrame.h.
  template <class KeyFrameT>
  void insertKeyFrame(KeyFrameT* kf, cv::Mat& color, cv::Mat& depth, cv::Mat& imgRGB) {
    cv::Mat T = kf->GetPose();
    float Tcw[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) Tcw[r * 4 + c] = T.template at<float>(r, c);
    insertKeyFrame(Tcw, kf->fx, kf->fy, kf->cx, kf->cy, color, depth, imgRGB);
  }
  void insertKeyFrame(const float Tcw[16], float fx, float fy, float cx, float cy, cv::Mat& /*gray*/, cv::Mat& depth,
                      cv::Mat& imgRGB) {
    if (ocm_insert_keyframe(h_, reinterpret_cast<const float*>(depth.data), imgRGB.data, depth.rows, depth.cols, Tcw, fx,
                            fy, cx, cy, nullptr) != B200ORB_OK)
      throw std::runtime_error(std::string("PointCloudMapping(B200): ") + b200orb_last_error());
  }
  // MapDrawer::UpdateOctomap (perfect/src/MapDrawer.cc:610-638): given the list of all keyframes so far, insert
  // [lastKeyframeSize, N-1) -- the reference never inserts the NEWEST keyframe (`i < N-1`, :615) -- and remember N-1.
  // KeyFrameT supplies GetPose(), fx/fy/cx/cy and the images the P variant stores on the keyframe (mImDep, mImRGB).
  template <class KeyFrameT>
  void UpdateOctomap(const std::vector<KeyFrameT*>& vKFs) {
    const int N = (int)vKFs.size();
    if (N > 1) {
      for (size_t i = lastKeyframeSize; i < (unsigned int)N - 1; i++) {
        cv::Mat gray;
        insertKeyFrame(vKFs[i], gray, vKFs[i]->mImDep, vKFs[i]->mImRGB);
      }
      lastKeyframeSize = N - 1;
    }
  }
  void shutdown() { ocm_sync(h_); }     // src/pointcloudmapping.cc:104-113 joins the viewer thread
  void update() {}                      // T variant's map refresh hook: nothing to do, the map lives in HBM
  long long numLeaves() { return (long long)ocm_num_leaves(h_); }
  ocm_t* handle() { return h_; }

 protected:
  double resolution;
  ocm_t* h_ = nullptr;
  size_t lastKeyframeSize = 0;
};

// The T variant of the same class (src/pointcloudmapping.cc): no occupancy tree, the map is the accumulated colour cloud
// `globalMap`, re-filtered by VoxelGrid(resolution) after every batch of new keyframes (:395-500).  insertKeyFrame()
// appends the keyframe's full cloud (generatePointCloud :131-194), update() runs the refilter the viewer thread runs when
// it wakes up, globalMapSize()/copyGlobalMap() expose what the viewer / global_color.pcd receive (:496-521).  The
// detector-driven semantic clusters of that file stay on the host.
class PointCloudMappingT {
 public:
  explicit PointCloudMappingT(double resolution_) : resolution(resolution_) {
    int dev = 0;
    if (const char* e = std::getenv("B200ORB_DEVICE")) dev = std::atoi(e);
    if (gcm_create((float)resolution_, dev, &h_) != B200ORB_OK)
      throw std::runtime_error(std::string("PointCloudMapping(B200, T): ") + b200orb_last_error());
  }
  ~PointCloudMappingT() { gcm_destroy(h_); }
  PointCloudMappingT(const PointCloudMappingT&) = delete;
  PointCloudMappingT& operator=(const PointCloudMappingT&) = delete;

  template <class KeyFrameT>
  void insertKeyFrame(KeyFrameT* kf, cv::Mat& color, cv::Mat& depth, cv::Mat& imgRGB) {
    cv::Mat T = kf->GetPose();
    float Tcw[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) Tcw[r * 4 + c] = T.template at<float>(r, c);
    insertKeyFrame(Tcw, kf->fx, kf->fy, kf->cx, kf->cy, color, depth, imgRGB);
  }
  void insertKeyFrame(const float Tcw[16], float fx, float fy, float cx, float cy, cv::Mat& /*gray*/, cv::Mat& depth,
                      cv::Mat& imgRGB) {
    if (gcm_add_keyframe(h_, reinterpret_cast<const float*>(depth.data), imgRGB.data, depth.rows, depth.cols, Tcw, fx, fy,
                         cx, cy) != B200ORB_OK)
      throw std::runtime_error(std::string("PointCloudMapping(B200, T): ") + b200orb_last_error());
  }
  void update() {   // the viewer thread's wake-up after new keyframes: voxel.filter(globalMap) + swap
    if (gcm_refilter(h_) != B200ORB_OK) throw std::runtime_error(std::string("PointCloudMapping(B200, T): ") + b200orb_last_error());
  }
  void shutdown() { gcm_sync(h_); }
  long long globalMapSize() { return gcm_size(h_); }
  long long copyGlobalMap(float* xyz, unsigned char* rgb, long long cap) {
    long long n = 0;
    if (gcm_export(h_, xyz, rgb, cap, &n) != B200ORB_OK) throw std::runtime_error(std::string("PointCloudMapping(B200, T): ") + b200orb_last_error());
    return n;
  }
  gcm_t* handle() { return h_; }

 protected:
  double resolution;
  gcm_t* h_ = nullptr;
};
#endif
