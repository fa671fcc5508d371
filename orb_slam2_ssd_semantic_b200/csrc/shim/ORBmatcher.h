// shim/ORBmatcher.h -- ORB_SLAM2::ORBmatcher surface (include/ORBmatcher.h:37-118) over the C-ABI.  The searches of
// the reference walk and MUTATE the Frame/MapPoint object graph; the shim flattens the graph into the array views of
// include/b200orb.h, calls the GPU, and applies the pointer mutations on the host in the reference's order.
// Written as templates over the reference's own Frame / MapPoint types so this header does not need Frame.h:
//   #include "Frame.h"  #include "MapPoint.h"  #include "shim/ORBmatcher.h"
//   namespace ORB_SLAM2 { typedef ORBmatcherT<Frame, MapPoint> ORBmatcher; }
// Implemented overloads: SearchByProjection(Frame&, const Frame&, th, bMono) (src/ORBmatcher.cc:1578-1724),
// SearchByProjection(Frame&, const std::vector<MapPoint*>&, th) (:63-156), SearchByBoW(KeyFrame*, Frame&, ...) (:217-363),
// SearchByBoW(KeyFrame*, KeyFrame*, ...) (:665-812); DescriptorDistance (:1968-1984).  The KeyFrame type is a template
// parameter of the BoW methods (its mFeatVec is any ordered map node id -> vector of keypoint indices, like
// DBoW2::FeatureVector).
#ifndef ORBMATCHER_SHIM_H
#define ORBMATCHER_SHIM_H

#include <stdexcept>
#include <string>
#include <vector>

#ifdef B200_SHIM_STANDIN
#include "cv_standin.h"
#else
#include <opencv2/opencv.hpp>
#endif

#include "../../../include/b200orb.h"

namespace ORB_SLAM2 {

template <class Frame, class MapPoint>
class ORBmatcherT {
 public:
  static const int TH_LOW = ORBM_TH_LOW, TH_HIGH = ORBM_TH_HIGH, HISTO_LENGTH = ORBM_HISTO_LENGTH;

  ORBmatcherT(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
    int dev = 0;
    if (const char* e = std::getenv("B200ORB_DEVICE")) dev = std::atoi(e);
    if (orbm_create(dev, &h_) != B200ORB_OK) throw std::runtime_error(std::string("ORBmatcher(B200): ") + b200orb_last_error());
  }
  ~ORBmatcherT() { orbm_destroy(h_); }

  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return orbm_hamming(a.ptr(0), b.ptr(0)); }

  // src/ORBmatcher.cc:1578-1724
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    FlatFrame cur;
    flatten(CurrentFrame, cur, /*with_mp_state=*/true);
    const int nl = LastFrame.N;
    std::vector<float> xw((size_t)nl * 3, 0.f), lang(nl);
    std::vector<uint8_t> valid(nl, 0), ldesc((size_t)nl * 32, 0);
    std::vector<int32_t> loct(nl), lobs(nl, 0);
    for (int i = 0; i < nl; ++i) {
      MapPoint* pMP = LastFrame.mvpMapPoints[i];
      loct[i] = LastFrame.mvKeys[i].octave;
      lang[i] = LastFrame.mvKeysUn[i].angle;
      if (pMP && !LastFrame.mvbOutlier[i]) {
        valid[i] = 1;
        cv::Mat p = pMP->GetWorldPos();
        for (int k = 0; k < 3; ++k) xw[(size_t)i * 3 + k] = p.template at<float>(k, 0);
        cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&ldesc[(size_t)i * 32], d.ptr(0), 32);
        lobs[i] = pMP->Observations();
      }
    }
    OrbmLast L;
    L.n = nl; L.xw = xw.data(); L.valid = valid.data(); L.octave = loct.data(); L.angle = lang.data();
    L.mp_desc = ldesc.data(); L.mp_obs = lobs.data();
    copy_pose(LastFrame.mTcw, L.Tcw);
    std::vector<int32_t> c2l(cur.n, -1);
    int nmatches = 0;
    if (orbm_search_by_projection_last(h_, &cur.view, &L, th, bMono ? 1 : 0, mfNNratio, mbCheckOrientation ? 1 : 0,
                                       c2l.data(), &nmatches) != B200ORB_OK)
      throw std::runtime_error(std::string("ORBmatcher(B200): ") + b200orb_last_error());
    for (int j = 0; j < cur.n; ++j) {      // pointer state the reference leaves behind (:1680,:1716)
      if (c2l[j] >= 0) CurrentFrame.mvpMapPoints[j] = LastFrame.mvpMapPoints[c2l[j]];
      else if (c2l[j] == -1) CurrentFrame.mvpMapPoints[j] = static_cast<MapPoint*>(NULL);
    }
    return nmatches;
  }

  // src/ORBmatcher.cc:63-156
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3) {
    FlatFrame cur;
    flatten(F, cur, true);
    const int np = (int)vpMapPoints.size();
    std::vector<uint8_t> inview(np, 0), desc((size_t)np * 32, 0);
    std::vector<float> px(np), py(np), pxr(np), vc(np);
    std::vector<int32_t> lvl(np, 0), obs(np, 0);
    for (int i = 0; i < np; ++i) {
      MapPoint* pMP = vpMapPoints[i];
      if (!pMP->mbTrackInView || pMP->isBad()) continue;
      inview[i] = 1;
      px[i] = pMP->mTrackProjX; py[i] = pMP->mTrackProjY; pxr[i] = pMP->mTrackProjXR;
      lvl[i] = pMP->mnTrackScaleLevel; vc[i] = pMP->mTrackViewCos; obs[i] = pMP->Observations();
      cv::Mat d = pMP->GetDescriptor();
      std::memcpy(&desc[(size_t)i * 32], d.ptr(0), 32);
    }
    OrbmTrackPoints P;
    P.n = np; P.track_in_view = inview.data(); P.proj_x = px.data(); P.proj_y = py.data(); P.proj_xr = pxr.data();
    P.scale_level = lvl.data(); P.view_cos = vc.data(); P.mp_desc = desc.data(); P.mp_obs = obs.data();
    std::vector<int32_t> f2p(cur.n, -1);
    int nmatches = 0;
    if (orbm_search_by_projection_points(h_, &cur.view, &P, th, mfNNratio, f2p.data(), &nmatches) != B200ORB_OK)
      throw std::runtime_error(std::string("ORBmatcher(B200): ") + b200orb_last_error());
    for (int j = 0; j < cur.n; ++j)
      if (f2p[j] >= 0) F.mvpMapPoints[j] = vpMapPoints[f2p[j]];
    return nmatches;
  }

  // src/ORBmatcher.cc:217-363: vpMapPointMatches[i] = MapPoint of the keyframe keypoint matched to frame keypoint i
  template <class KeyFrame>
  int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    FlatBow kf, fr;
    flatten_bow(pKF->mDescriptors, pKF->mvKeysUn, pKF->mFeatVec, &vpMapPointsKF, kf);
    flatten_bow(F.mDescriptors, F.mvKeys, F.mFeatVec, (const std::vector<MapPoint*>*)nullptr, fr);   // mvKeys: :308
    std::vector<int32_t> f2kf(fr.view.n > 0 ? fr.view.n : 1, -1);
    int nmatches = 0;
    if (orbm_search_by_bow(h_, &kf.view, &fr.view, mfNNratio, mbCheckOrientation ? 1 : 0, f2kf.data(), &nmatches) != B200ORB_OK)
      throw std::runtime_error(std::string("ORBmatcher(B200): ") + b200orb_last_error());
    vpMapPointMatches.assign((size_t)F.N, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < F.N; ++i)
      if (f2kf[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[f2kf[i]];
    return nmatches;
  }

  // src/ORBmatcher.cc:665-812: vpMatches12[i] = MapPoint of the pKF2 keypoint matched to pKF1 keypoint i
  template <class KeyFrame>
  int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    FlatBow k1, k2;
    flatten_bow(pKF1->mDescriptors, pKF1->mvKeysUn, pKF1->mFeatVec, &vpMapPoints1, k1);
    flatten_bow(pKF2->mDescriptors, pKF2->mvKeysUn, pKF2->mFeatVec, &vpMapPoints2, k2);
    std::vector<int32_t> m12(k1.view.n > 0 ? k1.view.n : 1, -1);
    int nmatches = 0;
    if (orbm_search_by_bow_kf(h_, &k1.view, &k2.view, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches) != B200ORB_OK)
      throw std::runtime_error(std::string("ORBmatcher(B200): ") + b200orb_last_error());
    vpMatches12.assign(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    for (size_t i = 0; i < vpMapPoints1.size(); ++i)
      if (m12[i] >= 0) vpMatches12[i] = vpMapPoints2[m12[i]];
    return nmatches;
  }

 protected:
  struct FlatBow {
    std::vector<uint8_t> desc, valid;
    std::vector<float> ang;
    std::vector<uint32_t> node_ids, idx;
    std::vector<int32_t> node_off;
    OrbmBow view;
  };
  // descriptors + angles + the FeatureVector (ordered map: ascending node ids) flattened; mps == NULL: no validity mask
  template <class Keys, class FeatVec>
  static void flatten_bow(const cv::Mat& descriptors, const Keys& keys, const FeatVec& fv, const std::vector<MapPoint*>* mps,
                          FlatBow& o) {
    const int n = (int)keys.size();
    o.desc.resize((size_t)(n > 0 ? n : 1) * 32);
    o.ang.resize(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) {
      std::memcpy(&o.desc[(size_t)i * 32], descriptors.ptr(i), 32);
      o.ang[i] = keys[i].angle;
    }
    if (mps) {
      o.valid.assign(n > 0 ? n : 1, 0);
      for (int i = 0; i < n; ++i) {
        MapPoint* p = (*mps)[i];
        o.valid[i] = (p && !p->isBad()) ? 1 : 0;
      }
    }
    o.node_off.assign(1, 0);
    for (typename FeatVec::const_iterator it = fv.begin(); it != fv.end(); ++it) {
      o.node_ids.push_back((uint32_t)it->first);
      for (size_t k = 0; k < it->second.size(); ++k) o.idx.push_back((uint32_t)it->second[k]);
      o.node_off.push_back((int32_t)o.idx.size());
    }
    if (o.idx.empty()) o.idx.push_back(0);
    if (o.node_ids.empty()) o.node_ids.push_back(0);
    OrbmBow& v = o.view;
    v.n = n; v.desc = o.desc.data(); v.angle = o.ang.data(); v.valid = mps ? o.valid.data() : nullptr;
    v.n_nodes = (int)o.node_off.size() - 1; v.node_ids = o.node_ids.data(); v.node_off = o.node_off.data(); v.idx = o.idx.data();
  }
  struct FlatFrame {
    int n = 0;
    std::vector<float> x, y, ang, ur, sf;
    std::vector<int32_t> oct, obs;
    std::vector<uint8_t> desc;
    OrbmFrame view;
  };
  static void copy_pose(const cv::Mat& T, float out[16]) {
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[r * 4 + c] = T.template at<float>(r, c);
  }
  static void flatten(Frame& F, FlatFrame& o, bool with_mp_state) {
    const int n = F.N;
    o.n = n;
    o.x.resize(n); o.y.resize(n); o.ang.resize(n); o.ur.resize(n); o.oct.resize(n); o.obs.assign(n, -1);
    o.desc.resize((size_t)n * 32);
    for (int i = 0; i < n; ++i) {
      o.x[i] = F.mvKeysUn[i].pt.x; o.y[i] = F.mvKeysUn[i].pt.y; o.ang[i] = F.mvKeysUn[i].angle;
      o.oct[i] = F.mvKeysUn[i].octave; o.ur[i] = F.mvuRight[i];
      std::memcpy(&o.desc[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
      if (with_mp_state && F.mvpMapPoints[i]) o.obs[i] = F.mvpMapPoints[i]->Observations();
    }
    o.sf.assign(F.mvScaleFactors.begin(), F.mvScaleFactors.end());
    OrbmFrame& v = o.view;
    v.n = n; v.x = o.x.data(); v.y = o.y.data(); v.octave = o.oct.data(); v.angle = o.ang.data(); v.uright = o.ur.data();
    v.desc = o.desc.data(); v.mp_obs = o.obs.data();
    copy_pose(F.mTcw, v.Tcw);
    v.fx = F.fx; v.fy = F.fy; v.cx = F.cx; v.cy = F.cy; v.bf = F.mbf; v.b = F.mb;
    v.min_x = Frame::mnMinX; v.max_x = Frame::mnMaxX; v.min_y = Frame::mnMinY; v.max_y = Frame::mnMaxY;
    v.scale_factors = o.sf.data(); v.nlevels = (int)o.sf.size();
  }
  float mfNNratio;
  bool mbCheckOrientation;
  orbm_t* h_ = nullptr;
};

}  // namespace ORB_SLAM2
#endif
