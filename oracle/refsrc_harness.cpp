// oracle/refsrc_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// C entry points around the REFERENCE'S OWN translation units, compiled unmodified from /root/reference into
// oracle/_ref/ (see oracle/Makefile, target `ref`): src/ORBextractor.cc, src/ORBmatcher.cc, src/Frame.cc,
// src/KeyFrame.cc, src/MapPoint.cc, src/Map.cc.  Each entry takes the same flat views as the restatement in
// oracle/match_ref.cpp / orb_ref.cpp (the C-ABI structs of include/b200orb.h), builds real ORB_SLAM2::Frame /
// KeyFrame / MapPoint objects from them, calls the real member function and flattens the pointer state it leaves.
// tests/test_oracle_cpu.py asserts restatement == this, bit for bit; that is what pins the oracle to reference code.
//
// Nothing here restates reference logic.  `#define private public` only opens the classes so that a Frame can be
// filled from arrays instead of from an image (the class layout is unchanged; the reference's .cc files are compiled
// without it).  OpenCV / DBoW2 are stand-ins (oracle/standin/), third-party arithmetic stated there.
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <new>
#include <set>
#include <sstream>
#include <thread>
#include <vector>

#include "../include/b200orb.h"
#include "refsrc_api.h"

#define private public
#define protected public
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "Map.h"
#include "KeyFrameDatabase.h"
#include "ORBVocabulary.h"
#undef private
#undef protected

namespace ORB_SLAM2 {
// the only out-of-path symbol the six translation units leave undefined (KeyFrame::SetBadFlag -> database erase)
void KeyFrameDatabase::erase(KeyFrame*) {}
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

// The same harness is compiled twice: against the reference's own ORBextractor / ORBmatcher (exports refsrc_*) and, with
// -DREFSRC_SHIM and the B200 shim headers shadowing include/ORBextractor.h and include/ORBmatcher.h, against the shims
// (exports shimsrc_*): the reference's Frame.cc / KeyFrame.cc / MapPoint.cc then run unchanged on top of libb200orb.so.
#ifdef REFSRC_SHIM
#define X(name) shimsrc_##name
#else
#define X(name) refsrc_##name
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Monotonic storage for the quad-tree's list nodes.  DistributeOctTree orders equal-size nodes by their HEAP ADDRESS
// (src/ORBextractor.cc:686 sorts pair<int, ExtractorNode*>), so the reference's keypoint order depends on the allocator
// of the build it runs in.  Here every std::list<ExtractorNode> node comes from a bump arena that
// never reuses memory inside one call, which makes "higher address" == "created later": the reference itself, made
// deterministic, and the rule the restatement and the kernels document ("later-created node first").
// Bound inside this library only (-Wl,-Bsymbolic); everything else goes to malloc/free.
namespace {
#ifdef REFSRC_SHIM
const size_t kNodeBytes = 0;   // no quad-tree on the host in the shim build
#else
const size_t kNodeBytes = sizeof(std::_List_node<ExtractorNode>);
#endif
const size_t kArenaBytes = 16u << 20;          // one per thread that extracts (a frame needs ~2 MB of list nodes)
const int kMaxArenas = 512;
// All arenas are slices of ONE reserved address range (mmap, MAP_NORESERVE: pages exist once touched), so operator delete
// recognises an arena pointer with two compares however many threads extract -- the frame-parallel CPU baseline runs
// nthreads extractors at once -- and a thread hands its slice back when it ends (the pipeline starts fresh threads per call).
char* g_region = nullptr;
std::once_flag g_region_once;
std::mutex g_free_mu;
std::vector<int> g_free;
int g_next = 0;
struct ArenaSlot {
  int id = -1;
  ~ArenaSlot() {
    if (id >= 0) { std::lock_guard<std::mutex> lk(g_free_mu); g_free.push_back(id); }
  }
};
thread_local ArenaSlot t_slot;
thread_local char* t_arena = nullptr;
thread_local size_t t_arena_used = 0;
thread_local bool t_arena_on = false;
inline void arena_reset() {
  if (!t_arena) {
    std::call_once(g_region_once, [] {
      void* r = mmap(nullptr, (size_t)kMaxArenas * kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      g_region = (r == MAP_FAILED) ? nullptr : (char*)r;
    });
    if (g_region) {
      std::lock_guard<std::mutex> lk(g_free_mu);
      int id = -1;
      if (!g_free.empty()) { id = g_free.back(); g_free.pop_back(); }
      else if (g_next < kMaxArenas) id = g_next++;
      if (id >= 0) { t_slot.id = id; t_arena = g_region + (size_t)id * kArenaBytes; }
    }
  }
  t_arena_used = 0;
  t_arena_on = t_arena != nullptr;
}
}  // namespace
void* operator new(size_t n) {
  if (t_arena_on && kNodeBytes && n == kNodeBytes && t_arena_used + ((n + 15) & ~size_t(15)) <= kArenaBytes) {
    void* p = t_arena + t_arena_used;
    t_arena_used += (n + 15) & ~size_t(15);
    return p;
  }
  void* p = malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void operator delete(void* p) noexcept {
  const char* r = g_region;
  if (r && (const char*)p >= r && (const char*)p < r + (size_t)kMaxArenas * kArenaBytes) return;
  free(p);
}
void operator delete(void* p, size_t) noexcept { operator delete(p); }

namespace {

std::mutex g_mu;   // Frame keeps intrinsics / bounds in statics: one harness call at a time
Map g_map;

cv::Mat mat44(const float* T) {
  cv::Mat m(4, 4, CV_32F);
  memcpy(m.data, T, 16 * sizeof(float));
  return m;
}
cv::Mat desc_row(const uint8_t* d) {
  cv::Mat m(1, 32, CV_8U);
  memcpy(m.data, d, 32);
  return m;
}

// A Frame filled from the flat view exactly as the RGB-D constructor leaves it (src/Frame.cc:176-240) minus the
// extraction itself: keypoints, descriptors, uRight, scale tables, statics, grid, pose.
Frame* make_frame(const OrbmFrame* f, const float* depth = nullptr) {
  Frame* F = new Frame();
  F->mpORBvocabulary = nullptr; F->mpORBextractorLeft = nullptr; F->mpORBextractorRight = nullptr;
  F->mnId = Frame::nNextId++;
  F->mTimeStamp = 0;
  F->mbf = f->bf; F->mb = f->b; F->mThDepth = 40.0f * f->b;
  F->N = f->n;
  F->mvKeys.resize(f->n); F->mvKeysUn.resize(f->n);
  for (int i = 0; i < f->n; ++i) {
    cv::KeyPoint kp(f->x[i], f->y[i], 31.f, f->angle[i], 0.f, f->octave[i]);
    F->mvKeys[i] = kp; F->mvKeysUn[i] = kp;
  }
  F->mvuRight.assign(f->uright, f->uright + f->n);
  F->mvDepth.assign(f->n, -1.f);
  if (depth) F->mvDepth.assign(depth, depth + f->n);
  F->mDescriptors.create(f->n, 32, CV_8U);
  if (f->n) memcpy(F->mDescriptors.data, f->desc, (size_t)f->n * 32);
  F->mvpMapPoints.assign(f->n, static_cast<MapPoint*>(NULL));
  F->mvbOutlier.assign(f->n, false);
  F->mnScaleLevels = f->nlevels;
  F->mfScaleFactor = f->nlevels > 1 ? f->scale_factors[1] : 1.2f;
  F->mfLogScaleFactor = log(F->mfScaleFactor);
  F->mvScaleFactors.assign(f->scale_factors, f->scale_factors + f->nlevels);
  F->mvInvScaleFactors.resize(f->nlevels); F->mvLevelSigma2.resize(f->nlevels); F->mvInvLevelSigma2.resize(f->nlevels);
  for (int l = 0; l < f->nlevels; ++l) {
    F->mvInvScaleFactors[l] = 1.0f / F->mvScaleFactors[l];
    F->mvLevelSigma2[l] = F->mvScaleFactors[l] * F->mvScaleFactors[l];
    F->mvInvLevelSigma2[l] = 1.0f / F->mvLevelSigma2[l];
  }
  Frame::fx = f->fx; Frame::fy = f->fy; Frame::cx = f->cx; Frame::cy = f->cy;
  Frame::invfx = 1.0f / f->fx; Frame::invfy = 1.0f / f->fy;
  Frame::mnMinX = f->min_x; Frame::mnMaxX = f->max_x; Frame::mnMinY = f->min_y; Frame::mnMaxY = f->max_y;
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
  Frame::mbInitialComputations = false;
  F->mK = cv::Mat::eye(3, 3, CV_32F);
  F->mK.at<float>(0, 0) = f->fx; F->mK.at<float>(1, 1) = f->fy; F->mK.at<float>(0, 2) = f->cx; F->mK.at<float>(1, 2) = f->cy;
  F->mDistCoef = cv::Mat::zeros(4, 1, CV_32F);
  F->AssignFeaturesToGrid();
  F->SetPose(mat44(f->Tcw));
  return F;
}

Frame* g_proto = nullptr;   // a one-keypoint frame MapPoints are constructed against
MapPoint* make_point(const float* xw, const uint8_t* desc, int obs) {
  if (!g_proto) {
    static float one = 1.f, zerof = 0.f, sf[8] = {1, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f, 3.5831808f};
    static int32_t zi = 0;
    static uint8_t zd[32] = {0};
    OrbmFrame p;
    memset(&p, 0, sizeof(p));
    p.n = 1; p.x = &one; p.y = &one; p.octave = &zi; p.angle = &zerof; p.uright = &zerof; p.desc = zd;
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(p.Tcw, I, sizeof(I));
    p.fx = p.fy = 500; p.cx = 320; p.cy = 240; p.bf = 40; p.b = 0.08f; p.min_x = 0; p.max_x = 640; p.min_y = 0; p.max_y = 480;
    p.scale_factors = sf; p.nlevels = 8;
    g_proto = make_frame(&p);
  }
  cv::Mat pos(3, 1, CV_32F);
  const float far[3] = {0.1f, 0.2f, 1.0f};
  memcpy(pos.data, xw ? xw : far, 12);
  MapPoint* mp = new MapPoint(pos, &g_map, g_proto, 0);
  if (desc) desc_row(desc).copyTo(mp->mDescriptor);
  mp->nObs = obs;
  return mp;
}

void restore_statics(const OrbmFrame* f) {   // make_point's proto frame overwrote them on first use
  Frame::fx = f->fx; Frame::fy = f->fy; Frame::cx = f->cx; Frame::cy = f->cy;
  Frame::invfx = 1.0f / f->fx; Frame::invfy = 1.0f / f->fy;
  Frame::mnMinX = f->min_x; Frame::mnMaxX = f->max_x; Frame::mnMinY = f->min_y; Frame::mnMaxY = f->max_y;
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
}

// pre-existing map points of the current frame (cur->mp_obs >= 0): reported as -2 while they survive
void seed_existing(Frame* F, const OrbmFrame* f, std::map<MapPoint*, int>& id) {
  if (!f->mp_obs) return;
  for (int j = 0; j < f->n; ++j)
    if (f->mp_obs[j] >= 0) { MapPoint* mp = make_point(nullptr, nullptr, f->mp_obs[j]); F->mvpMapPoints[j] = mp; id[mp] = -2; }
}
void flatten(const std::vector<MapPoint*>& v, const std::map<MapPoint*, int>& id, int32_t* out) {
  for (size_t j = 0; j < v.size(); ++j) out[j] = v[j] ? id.at(v[j]) : -1;
}
void set_featvec(DBoW2::FeatureVector& fv, const OrbmBow* b) {
  fv.clear();
  for (int k = 0; k < b->n_nodes; ++k)
    for (int j = b->node_off[k]; j < b->node_off[k + 1]; ++j) fv.addFeature(b->node_ids[k], b->idx[j]);
}
OrbmFrame bow_as_frame(const OrbmBow* b, std::vector<float>& zeros, std::vector<int32_t>& zi) {
  static float sf[8] = {1, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f, 3.5831808f};
  zeros.assign(b->n, 0.f); zi.assign(b->n, 0);
  OrbmFrame p;
  memset(&p, 0, sizeof(p));
  p.n = b->n; p.x = zeros.data(); p.y = zeros.data(); p.octave = zi.data(); p.angle = b->angle; p.uright = zeros.data();
  p.desc = b->desc;
  const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memcpy(p.Tcw, I, sizeof(I));
  p.fx = p.fy = 500; p.cx = 320; p.cy = 240; p.bf = 40; p.b = 0.08f; p.min_x = 0; p.max_x = 640; p.min_y = 0; p.max_y = 480;
  p.scale_factors = sf; p.nlevels = 8;
  return p;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------------------------
// ORBextractor (src/ORBextractor.cc, whole file)
void* X(orb_create)(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  return new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void X(orb_destroy)(void* h) { delete (ORBextractor*)h; }
int X(orb_extract)(void* h, const uint8_t* img, int rows, int cols, int stride, void* kps, uint8_t* desc, int cap,
                       int* n_out) {
  ORBextractor* e = (ORBextractor*)h;
  std::lock_guard<std::mutex> lk(g_mu);
  arena_reset();
  cv::Mat image(rows, cols, CV_8UC1, (void*)img, (size_t)stride);
  std::vector<cv::KeyPoint> keys;
  cv::Mat descriptors;
  (*e)(image, cv::Mat(), keys, descriptors);
  *n_out = (int)keys.size();
  if ((int)keys.size() > cap) return -2;
  static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
  if (!keys.empty()) {
    memcpy(kps, keys.data(), keys.size() * sizeof(cv::KeyPoint));
    for (size_t i = 0; i < keys.size(); ++i) memcpy(desc + 32 * i, descriptors.ptr((int)i), 32);
  }
  return 0;
}
#ifndef REFSRC_SHIM   // internals of the reference's extractor class (the shim has its own, C-ABI backed, members)
void X(orb_tables)(void* h, float* sf, float* invsf, float* sigma2, float* invsigma2, int* nfeat, int* umax) {
  ORBextractor* e = (ORBextractor*)h;
  for (int l = 0; l < e->nlevels; ++l) {
    sf[l] = e->mvScaleFactor[l]; invsf[l] = e->mvInvScaleFactor[l]; sigma2[l] = e->mvLevelSigma2[l];
    invsigma2[l] = e->mvInvLevelSigma2[l]; nfeat[l] = e->mnFeaturesPerLevel[l];
  }
  for (int v = 0; v < 16; ++v) umax[v] = e->umax[v];
}
int X(orb_level_dims)(void* h, int level, int* w, int* hgt) {
  ORBextractor* e = (ORBextractor*)h;
  if (level < 0 || level >= e->nlevels || e->mvImagePyramid[level].empty()) return -1;
  *w = e->mvImagePyramid[level].cols; *hgt = e->mvImagePyramid[level].rows;
  return 0;
}
// mvImagePyramid[level]; bordered=1 returns the (w+38)x(h+38) parent buffer the ROI lives in (src/ORBextractor.cc:1125-1129)
void X(orb_get_level)(void* h, int level, int bordered, uint8_t* dst) {
  ORBextractor* e = (ORBextractor*)h;
  const cv::Mat& m = e->mvImagePyramid[level];
  const int B = bordered ? 19 : 0;
  const int W = m.cols + 2 * B, H = m.rows + 2 * B;
  const uint8_t* base = m.data - (size_t)B * m.step - B;
  for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * W, base + (size_t)y * m.step, (size_t)W);
}
// ORBextractor::DistributeOctTree (src/ORBextractor.cc:540-765) on a caller-supplied candidate list
int X(orb_distribute)(const void* in, int n_in, int minX, int maxX, int minY, int maxY, int N, void* out, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  arena_reset();
  ORBextractor e(std::max(N, 1), 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> v((const cv::KeyPoint*)in, (const cv::KeyPoint*)in + n_in);
  std::vector<cv::KeyPoint> r = e.DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
  if ((int)r.size() > cap) return -2;
  if (!r.empty()) memcpy(out, r.data(), r.size() * sizeof(cv::KeyPoint));
  return (int)r.size();
}

#endif   // !REFSRC_SHIM

// ---------------------------------------------------------------------------------------------------------------------
// ORBmatcher (src/ORBmatcher.cc) -- same signatures as match_ref_* in oracle/match_ref.cpp
int X(hamming)(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

// SearchByProjection(Frame&, const Frame&, th, bMono), src/ORBmatcher.cc:1578-1724
int X(projection_last)(const OrbmFrame* cur, const OrbmLast* last, float th, int mono, float nnratio, int check_ori,
                           int32_t* cur2last, int* nmatches_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<MapPoint*, int> id;
  make_point(nullptr, nullptr, 0);   // force the proto frame before the statics are set for `cur`
  // the last frame: only N, mvpMapPoints, mvbOutlier, mvKeys[i].octave, mvKeysUn[i].angle and mTcw are read
  std::vector<float> zf(last->n, 0.f);
  std::vector<uint8_t> zd((size_t)std::max(last->n, 1) * 32, 0);
  OrbmFrame lv = *cur;
  lv.n = last->n; lv.x = zf.data(); lv.y = zf.data(); lv.octave = last->octave; lv.angle = last->angle; lv.uright = zf.data();
  lv.desc = zd.data(); lv.mp_obs = nullptr;
  memcpy(lv.Tcw, last->Tcw, sizeof(lv.Tcw));
  Frame* L = make_frame(&lv);
  for (int i = 0; i < last->n; ++i)
    if (last->valid[i]) {
      MapPoint* mp = make_point(last->xw + 3 * i, last->mp_desc + 32 * (size_t)i, last->mp_obs ? last->mp_obs[i] : 0);
      L->mvpMapPoints[i] = mp; id[mp] = i;
    }
  Frame* C = make_frame(cur);
  seed_existing(C, cur, id);
  restore_statics(cur);
  ORBmatcher m(nnratio, check_ori != 0);
  *nmatches_out = m.SearchByProjection(*C, *L, th, mono != 0);
  flatten(C->mvpMapPoints, id, cur2last);
  for (auto& kv : id) delete kv.first;
  delete C; delete L;
  return 0;
}

// SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:63-156
int X(projection_points)(const OrbmFrame* F, const OrbmTrackPoints* pts, float th, float nnratio, int32_t* f2pt,
                             int* nmatches_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<MapPoint*, int> id;
  make_point(nullptr, nullptr, 0);
  std::vector<MapPoint*> v(pts->n);
  for (int i = 0; i < pts->n; ++i) {
    MapPoint* mp = make_point(nullptr, pts->mp_desc + 32 * (size_t)i, pts->mp_obs ? pts->mp_obs[i] : 1);
    mp->mbTrackInView = pts->track_in_view[i] != 0;
    mp->mTrackProjX = pts->proj_x[i]; mp->mTrackProjY = pts->proj_y[i]; mp->mTrackProjXR = pts->proj_xr[i];
    mp->mnTrackScaleLevel = pts->scale_level[i]; mp->mTrackViewCos = pts->view_cos[i];
    v[i] = mp; id[mp] = i;
  }
  Frame* C = make_frame(F);
  seed_existing(C, F, id);
  restore_statics(F);
  ORBmatcher m(nnratio, true);
  *nmatches_out = m.SearchByProjection(*C, v, th);
  flatten(C->mvpMapPoints, id, f2pt);
  for (auto& kv : id) delete kv.first;
  delete C;
  return 0;
}

// SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), src/ORBmatcher.cc:217-363
int X(bow)(const OrbmBow* kf, const OrbmBow* f, float nnratio, int check_ori, int32_t* f2kf, int* nmatches_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<MapPoint*, int> id;
  make_point(nullptr, nullptr, 0);
  std::vector<float> z1, z2; std::vector<int32_t> i1, i2;
  OrbmFrame v1 = bow_as_frame(kf, z1, i1), v2 = bow_as_frame(f, z2, i2);
  Frame* FK = make_frame(&v1);
  for (int i = 0; i < kf->n; ++i)
    if (kf->valid[i]) { MapPoint* mp = make_point(nullptr, nullptr, 1); FK->mvpMapPoints[i] = mp; id[mp] = i; }
  set_featvec(FK->mFeatVec, kf);
  KeyFrame* K = new KeyFrame(*FK, &g_map, nullptr);
  Frame* F = make_frame(&v2);
  set_featvec(F->mFeatVec, f);
  std::vector<MapPoint*> out;
  ORBmatcher m(nnratio, check_ori != 0);
  *nmatches_out = m.SearchByBoW(K, *F, out);
  flatten(out, id, f2kf);
  for (auto& kv : id) delete kv.first;
  delete K; delete F; delete FK;
  return 0;
}

// SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&), src/ORBmatcher.cc:665-812
int X(bow_kf)(const OrbmBow* k1, const OrbmBow* k2, float nnratio, int check_ori, int32_t* matches12, int* nmatches_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<MapPoint*, int> id;
  make_point(nullptr, nullptr, 0);
  std::vector<float> z1, z2; std::vector<int32_t> i1, i2;
  OrbmFrame v1 = bow_as_frame(k1, z1, i1), v2 = bow_as_frame(k2, z2, i2);
  Frame* F1 = make_frame(&v1);
  Frame* F2 = make_frame(&v2);
  std::vector<MapPoint*> own;
  for (int i = 0; i < k1->n; ++i)
    if (k1->valid[i]) { MapPoint* mp = make_point(nullptr, nullptr, 1); F1->mvpMapPoints[i] = mp; own.push_back(mp); }
  for (int i = 0; i < k2->n; ++i)
    if (k2->valid[i]) { MapPoint* mp = make_point(nullptr, nullptr, 1); F2->mvpMapPoints[i] = mp; id[mp] = i; }
  set_featvec(F1->mFeatVec, k1);
  set_featvec(F2->mFeatVec, k2);
  KeyFrame* K1 = new KeyFrame(*F1, &g_map, nullptr);
  KeyFrame* K2 = new KeyFrame(*F2, &g_map, nullptr);
  std::vector<MapPoint*> out;
  ORBmatcher m(nnratio, check_ori != 0);
  *nmatches_out = m.SearchByBoW(K1, K2, out);
  flatten(out, id, matches12);
  for (auto& kv : id) delete kv.first;
  for (auto p : own) delete p;
  delete K1; delete K2; delete F1; delete F2;
  return 0;
}

// SearchForInitialization(Frame&, Frame&, vector<Point2f>&, vector<int>&, windowSize), src/ORBmatcher.cc:523-651
int X(initialization)(const OrbmFrame* f1, const OrbmFrame* f2, float* prev_xy, int window, float nnratio, int check_ori,
                          int32_t* matches12, int* nmatches_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  Frame* F1 = make_frame(f1);
  Frame* F2 = make_frame(f2);
  std::vector<cv::Point2f> prev(f1->n);
  for (int i = 0; i < f1->n; ++i) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher m(nnratio, check_ori != 0);
  *nmatches_out = m.SearchForInitialization(*F1, *F2, prev, m12, window);
  for (int i = 0; i < f1->n; ++i) { matches12[i] = m12[i]; prev_xy[2 * i] = prev[i].x; prev_xy[2 * i + 1] = prev[i].y; }
  delete F1; delete F2;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Frame glue through the REAL RGB-D constructor (src/Frame.cc:176-240): ExtractORB, UndistortKeyPoints,
// ComputeStereoFromRGBD (:850-871), AssignFeaturesToGrid, then UnprojectStereo (:879-899) per keypoint.
// depth f32 metres.  Outputs per keypoint: cv::KeyPoint (mvKeysUn), descriptor, uRight, depth, world point, valid.
int X(frame_rgbd)(const uint8_t* gray, const float* depth, int rows, int cols, int nfeatures, float scaleFactor,
                      int nlevels, int iniTh, int minTh, const float* Tcw, float fx, float fy, float cx, float cy,
                      float bf, const float* dist4, void* kps_un, uint8_t* desc, float* uright, float* kdepth, float* xw,
                      uint8_t* valid, int cap, int* n_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  arena_reset();
  ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
  K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy;
  cv::Mat D(4, 1, CV_32F);
  for (int i = 0; i < 4; ++i) D.at<float>(i) = dist4 ? dist4[i] : 0.f;
  cv::Mat g(rows, cols, CV_8UC1, (void*)gray), d(rows, cols, CV_32F, (void*)depth);
  Frame::mbInitialComputations = true;
  const float thDepth = bf * 40.0f / fx;
  Frame F(g, d, 0.0, &ex, nullptr, K, D, bf, thDepth);
  F.SetPose(mat44(Tcw));
  *n_out = F.N;
  if (F.N > cap) return -2;
  for (int i = 0; i < F.N; ++i) {
    memcpy((char*)kps_un + 28 * (size_t)i, &F.mvKeysUn[i], 28);
    memcpy(desc + 32 * (size_t)i, F.mDescriptors.ptr(i), 32);
    uright[i] = F.mvuRight[i]; kdepth[i] = F.mvDepth[i];
    cv::Mat x = F.UnprojectStereo(i);
    valid[i] = !x.empty();
    for (int k = 0; k < 3; ++k) xw[3 * i + k] = x.empty() ? 0.f : x.at<float>(k);
  }
  return 0;
}

// CPU baseline through the reference's own tracking sources: per frame the RGB-D Frame constructor (ORBextractor,
// ComputeStereoFromRGBD, AssignFeaturesToGrid), then per consecutive pair the map points Tracking would hold for the last
// frame (a MapPoint per keypoint with depth, src/Tracking.cc:1285-1322 UpdateLastFrame style) and
// ORBmatcher::SearchByProjection(cur, last, th, false).  Frame-parallel on `nthreads` threads (one ORBextractor per thread,
// like src/Frame.cc:121-124); Frame's statics are written once before the threads start.  depth f32 metres.
// Returns seconds (steady_clock, like perfect/Examples/RGB-D/rgbd_tum.cc:92-111).
double X(pipeline_run)(const uint8_t* gray, const float* depth, const float* Tcw, int n, int rows, int cols, int nfeatures,
                           float scaleFactor, int nlevels, int iniTh, int minTh, float fx, float fy, float cx, float cy,
                           float bf, float th, float nnratio, int check_ori, int nthreads, int* nkp, int* nmatch) {
  std::lock_guard<std::mutex> lk(g_mu);
  make_point(nullptr, nullptr, 0);
  cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
  K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy;
  cv::Mat D = cv::Mat::zeros(4, 1, CV_32F);
  const float thDepth = bf * 40.0f / fx;
  const size_t px = (size_t)rows * cols;
  {   // first frame constructed once up front: it performs Frame's one-off static initialisation (src/Frame.cc:214-231)
    ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
    Frame::mbInitialComputations = true;
    arena_reset();
    cv::Mat g(rows, cols, CV_8UC1, (void*)gray), d(rows, cols, CV_32F, (void*)depth);
    Frame F(g, d, 0.0, &ex, nullptr, K, D, bf, thDepth);
  }
  std::vector<Frame*> frames(n, nullptr);
  auto t0 = std::chrono::steady_clock::now();
  {
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&]() {
        ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        cv::Mat Kt = K.clone(), Dt = D.clone();
        for (int i = next++; i < n; i = next++) {
          arena_reset();
          cv::Mat g(rows, cols, CV_8UC1, (void*)(gray + px * i)), d(rows, cols, CV_32F, (void*)(depth + px * i));
          Frame* F = new Frame(g, d, (double)i, &ex, nullptr, Kt, Dt, bf, thDepth);
          F->mpORBextractorLeft = nullptr;
          F->SetPose(mat44(Tcw + 16 * i));
          frames[i] = F;
          nkp[i] = F->N;
        }
      });
    for (auto& t : pool) t.join();
  }
  nmatch[0] = 0;
  {
    std::atomic<int> next(1);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&]() {
        Map tmap;   // the reference has ONE tracking thread creating map points; here every worker plays that thread
        for (int i = next++; i < n; i = next++) {
          Frame last(*frames[i - 1]);
          Frame cur(*frames[i]);
          std::vector<MapPoint*> mine;
          for (int k = 0; k < last.N; ++k) {
            if (last.mvDepth[k] <= 0) continue;
            cv::Mat x3D = last.UnprojectStereo(k);
            MapPoint* mp = new MapPoint(x3D, &tmap, &last, k);
            mp->nObs = 1;
            last.mvpMapPoints[k] = mp;
            mine.push_back(mp);
          }
          ORBmatcher m(nnratio, check_ori != 0);
          nmatch[i] = m.SearchByProjection(cur, last, th, false);
          for (MapPoint* p : mine) delete p;
        }
      });
    for (auto& t : pool) t.join();
  }
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (Frame* F : frames) delete F;
  return sec;
}

// Frame::isInFrustum (src/Frame.cc:387-451) for a list of world points with given normal / distance bounds.
// out per point: in_view, proj_x, proj_y, proj_xr, scale_level, view_cos
int X(is_in_frustum)(const OrbmFrame* f, int n, const float* xw, const float* normal, const float* min_dist,
                         const float* max_dist, float viewing_cos_limit, uint8_t* in_view, float* proj_x, float* proj_y,
                         float* proj_xr, int32_t* scale_level, float* view_cos) {
  std::lock_guard<std::mutex> lk(g_mu);
  make_point(nullptr, nullptr, 0);
  Frame* F = make_frame(f);
  restore_statics(f);
  for (int i = 0; i < n; ++i) {
    MapPoint* mp = make_point(xw + 3 * i, nullptr, 1);
    memcpy(mp->mNormalVector.data, normal + 3 * i, 12);
    mp->mfMinDistance = min_dist[i]; mp->mfMaxDistance = max_dist[i];   // raw mfMinDistance / mfMaxDistance
    mp->mTrackProjX = mp->mTrackProjY = mp->mTrackProjXR = 0; mp->mnTrackScaleLevel = 0; mp->mTrackViewCos = 0;
    in_view[i] = F->isInFrustum(mp, viewing_cos_limit);
    proj_x[i] = mp->mTrackProjX; proj_y[i] = mp->mTrackProjY; proj_xr[i] = mp->mTrackProjXR;
    scale_level[i] = mp->mnTrackScaleLevel; view_cos[i] = mp->mTrackViewCos;
    delete mp;
  }
  delete F;
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// The remaining ORBmatcher members on real KeyFrame / MapPoint graphs
namespace {
struct Points {
  std::vector<MapPoint*> v;
  std::map<MapPoint*, int> id;
  ~Points() { for (auto p : v) delete p; }
};
void make_points(const RefMapPoints* r, Points& out, int id_base = 0) {
  out.v.assign(r->n, static_cast<MapPoint*>(NULL));
  for (int i = 0; i < r->n; ++i) {
    if (r->valid && !r->valid[i]) continue;
    MapPoint* mp = make_point(r->xw + 3 * i, r->desc ? r->desc + 32 * (size_t)i : nullptr, r->obs ? r->obs[i] : 1);
    if (r->normal) memcpy(mp->mNormalVector.data, r->normal + 3 * i, 12);
    mp->mfMinDistance = r->min_dist ? r->min_dist[i] : 0.f;
    mp->mfMaxDistance = r->max_dist ? r->max_dist[i] : 1e9f;
    if (r->bad && r->bad[i]) mp->mbBad = true;
    out.v[i] = mp; out.id[mp] = id_base + i;
  }
}
// KeyFrame from a flat frame view whose keypoint j holds MapPoint kfmp.v[j] (registered as an observation so that
// Replace / IsInKeyFrame / GetIndexInKeyFrame work on the real graph)
KeyFrame* make_keyframe(const OrbmFrame* f, Points* kfmp, Frame** keep, const float* depth = nullptr) {
  Frame* F = make_frame(f, depth);
  KeyFrame* K = new KeyFrame(*F, &g_map, nullptr);
  if (kfmp)
    for (int j = 0; j < f->n && j < (int)kfmp->v.size(); ++j)
      if (kfmp->v[j]) {
        K->AddMapPoint(kfmp->v[j], j);
        const int keep_obs = kfmp->v[j]->nObs;
        kfmp->v[j]->AddObservation(K, j);
        kfmp->v[j]->nObs = keep_obs;          // Observations() as the view states it
      }
  *keep = F;
  return K;
}
}  // namespace

// SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist), src/ORBmatcher.cc:1757-1899 (relocalisation).
// kf_mps: one entry per keyframe keypoint; already_found[i] != 0 puts that MapPoint into sAlreadyFound.
int X(projection_kf)(const OrbmFrame* cur, const OrbmFrame* kf, const RefMapPoints* kf_mps, const uint8_t* already_found,
                     float th, int orb_dist, float nnratio, int check_ori, int32_t* cur2kf, int* nmatches_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  make_point(nullptr, nullptr, 0);
  Points P;
  make_points(kf_mps, P);
  Frame* FK;
  KeyFrame* K = make_keyframe(kf, &P, &FK);
  Frame* C = make_frame(cur);
  Points pre;
  std::map<MapPoint*, int> id = P.id;
  seed_existing(C, cur, id);
  restore_statics(cur);
  std::set<MapPoint*> found;
  for (int i = 0; i < kf_mps->n; ++i)
    if (already_found && already_found[i] && P.v[i]) found.insert(P.v[i]);
  ORBmatcher m(nnratio, check_ori != 0);
  *nmatches_out = m.SearchByProjection(*C, K, found, th, orb_dist);
  flatten(C->mvpMapPoints, id, cur2kf);
  for (auto& kv : id) if (!P.id.count(kv.first)) delete kv.first;
  delete C; delete K; delete FK;
  return 0;
}

// SearchByProjection(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, vector<MapPoint*>& vpMatched, int th),
// src/ORBmatcher.cc:378-498 (loop closing).  matched_io[kf->n]: index into pts of the MapPoint already matched to the
// keypoint, -1 none, -2 some MapPoint that is not in pts; updated in place.
int X(projection_sim3)(const OrbmFrame* kf, const float* Scw, const RefMapPoints* pts, int32_t* matched_io, int th,
                       int* nmatches_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  make_point(nullptr, nullptr, 0);
  Points P;
  make_points(pts, P);
  for (int i = 0; i < pts->n; ++i)   // the reference dereferences every entry (:408-409): a missing point becomes a bad one
    if (!P.v[i]) { P.v[i] = make_point(nullptr, nullptr, 1); P.v[i]->mbBad = true; P.id[P.v[i]] = i; }
  Frame* FK;
  KeyFrame* K = make_keyframe(kf, nullptr, &FK);
  restore_statics(kf);
  std::vector<MapPoint*> vpMatched(kf->n, static_cast<MapPoint*>(NULL));
  std::map<MapPoint*, int> id = P.id;
  std::vector<MapPoint*> others;
  for (int j = 0; j < kf->n; ++j) {
    if (matched_io[j] >= 0) vpMatched[j] = P.v[matched_io[j]];
    else if (matched_io[j] == -2) { MapPoint* o = make_point(nullptr, nullptr, 1); others.push_back(o); id[o] = -2; vpMatched[j] = o; }
  }
  ORBmatcher m(0.75f, true);
  *nmatches_out = m.SearchByProjection(K, mat44(Scw), P.v, vpMatched, th);
  flatten(vpMatched, id, matched_io);
  for (auto o : others) delete o;
  delete K; delete FK;
  return 0;
}

// Fuse(KeyFrame*, const vector<MapPoint*>&, th), src/ORBmatcher.cc:1031-1182.  kf_mps: the keyframe's own MapPoints (one
// per keypoint, valid = 0 where none); pts: the candidates; in_kf[i] != 0 makes candidate i an observation of the keyframe
// already (IsInKeyFrame).  Outputs: slot_owner[kf->n] = who sits on each keypoint afterwards (-1 nobody, i = candidate i,
// 1000000 + j = the keyframe's original MapPoint j), replaced_by[pts->n] = same coding for GetReplaced() of candidate i
// (-1: not replaced), kf_replaced_by[kf->n] likewise for the original MapPoints.
int X(fuse)(const OrbmFrame* kf, const RefMapPoints* kf_mps, const RefMapPoints* pts, const uint8_t* in_kf, float th,
            const float* depth, int32_t* slot_owner, int32_t* replaced_by, int32_t* kf_replaced_by, int* nfused_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  make_point(nullptr, nullptr, 0);
  Points KP, P;
  make_points(kf_mps, KP, 1000000);
  make_points(pts, P, 0);
  Frame* FK;
  KeyFrame* K = make_keyframe(kf, &KP, &FK, depth);
  restore_statics(kf);
  // a far-away second keyframe holds the candidates flagged in_kf?  No: IsInKeyFrame(pKF) must be true for THIS keyframe:
  // register the observation on a keypoint slot that holds nobody (the flag only has to exist)
  for (int i = 0; i < pts->n; ++i)
    if (in_kf && in_kf[i] && P.v[i]) { const int keep = P.v[i]->nObs; P.v[i]->mObservations[K] = (size_t)0; P.v[i]->nObs = keep; }
  ORBmatcher m(0.6f, true);
  *nfused_out = m.Fuse(K, P.v, th);
  std::map<MapPoint*, int> id = P.id;
  id.insert(KP.id.begin(), KP.id.end());
  const std::vector<MapPoint*> now = K->GetMapPointMatches();
  for (int j = 0; j < kf->n; ++j) slot_owner[j] = now[j] ? id.at(now[j]) : -1;
  for (int i = 0; i < pts->n; ++i) { MapPoint* r = P.v[i] ? P.v[i]->GetReplaced() : nullptr; replaced_by[i] = r ? id.at(r) : -1; }
  for (int j = 0; j < kf->n; ++j) { MapPoint* r = KP.v[j] ? KP.v[j]->GetReplaced() : nullptr; kf_replaced_by[j] = r ? id.at(r) : -1; }
  delete K; delete FK;
  return 0;
}

// Fuse(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, float th, vector<MapPoint*>& vpReplacePoint),
// src/ORBmatcher.cc:1198-1318.  replace_out[pts->n]: 1000000 + j when vpReplacePoint[i] is the keyframe's MapPoint j.
int X(fuse_sim3)(const OrbmFrame* kf, const RefMapPoints* kf_mps, const float* Scw, const RefMapPoints* pts, float th,
                 int32_t* slot_owner, int32_t* replace_out, int* nfused_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  make_point(nullptr, nullptr, 0);
  Points KP, P;
  make_points(kf_mps, KP, 1000000);
  make_points(pts, P, 0);
  Frame* FK;
  KeyFrame* K = make_keyframe(kf, &KP, &FK);
  restore_statics(kf);
  std::vector<MapPoint*> rep(pts->n, static_cast<MapPoint*>(NULL));
  ORBmatcher m(0.8f, true);
  *nfused_out = m.Fuse(K, mat44(Scw), P.v, th, rep);
  std::map<MapPoint*, int> id = P.id;
  id.insert(KP.id.begin(), KP.id.end());
  const std::vector<MapPoint*> now = K->GetMapPointMatches();
  for (int j = 0; j < kf->n; ++j) slot_owner[j] = now[j] ? id.at(now[j]) : -1;
  for (int i = 0; i < pts->n; ++i) replace_out[i] = rep[i] ? id.at(rep[i]) : -1;
  delete K; delete FK;
  return 0;
}

// SearchBySim3(KeyFrame*, KeyFrame*, vector<MapPoint*>& vpMatches12, s12, R12, t12, th), src/ORBmatcher.cc:1334-1558.
// mp1 / mp2: the keyframes' MapPoints (one per keypoint); matches12_io[kf1->n]: pKF2 keypoint whose MapPoint is already
// matched to pKF1 keypoint i (-1 none), updated in place with the new matches.
int X(search_by_sim3)(const OrbmFrame* kf1, const OrbmFrame* kf2, const RefMapPoints* mp1, const RefMapPoints* mp2,
                      int32_t* matches12_io, float s12, const float* R12, const float* t12, float th, int* nfound_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  make_point(nullptr, nullptr, 0);
  Points P1, P2;
  make_points(mp1, P1, 1000000);
  make_points(mp2, P2, 0);
  Frame *F1, *F2;
  KeyFrame* K1 = make_keyframe(kf1, &P1, &F1);
  KeyFrame* K2 = make_keyframe(kf2, &P2, &F2);
  restore_statics(kf1);
  std::vector<MapPoint*> v12(kf1->n, static_cast<MapPoint*>(NULL));
  for (int i = 0; i < kf1->n; ++i)
    if (matches12_io[i] >= 0) v12[i] = P2.v[matches12_io[i]];
  cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F);
  memcpy(R.data, R12, 36); memcpy(t.data, t12, 12);
  ORBmatcher m(0.75f, true);
  *nfound_out = m.SearchBySim3(K1, K2, v12, s12, R, t, th);
  flatten(v12, P2.id, matches12_io);
  delete K1; delete K2; delete F1; delete F2;
  return 0;
}

// SearchForTriangulation(KeyFrame*, KeyFrame*, cv::Mat F12, vector<pair<size_t,size_t>>&, bOnlyStereo),
// src/ORBmatcher.cc:827-1019.  k1 / k2 carry keypoints, right coordinates, MapPoint occupancy and the FeatureVector;
// Tcw1 / Tcw2 the poses (for the epipole).  epipole_out[2] = (ex, ey) as the function computes it.
int X(triangulation)(const OrbmTriKF* k1, const OrbmTriKF* k2, const float* Tcw1, const float* Tcw2, const OrbmFrame* cam,
                     const float* F12, int only_stereo, float nnratio, int check_ori, int32_t* matches12, int* nmatches_out,
                     float* epipole_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  make_point(nullptr, nullptr, 0);
  auto build = [&](const OrbmTriKF* k, const float* Tcw, Frame** keep, Points& P) {
    OrbmFrame v = *cam;
    v.n = k->n; v.x = k->x; v.y = k->y; v.octave = k->octave; v.angle = k->angle; v.uright = k->uright; v.desc = k->desc;
    v.mp_obs = nullptr;
    memcpy(v.Tcw, Tcw, 64);
    Frame* F = make_frame(&v);
    OrbmBow b;
    memset(&b, 0, sizeof(b));
    b.n_nodes = k->n_nodes; b.node_ids = k->node_ids; b.node_off = k->node_off; b.idx = k->idx;
    set_featvec(F->mFeatVec, &b);
    P.v.assign(k->n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < k->n; ++i)
      if (k->has_mp && k->has_mp[i]) P.v[i] = make_point(nullptr, nullptr, 1);
    KeyFrame* K = new KeyFrame(*F, &g_map, nullptr);
    for (int i = 0; i < k->n; ++i) if (P.v[i]) K->AddMapPoint(P.v[i], i);
    *keep = F;
    return K;
  };
  Points P1, P2;
  Frame *F1, *F2;
  KeyFrame* K1 = build(k1, Tcw1, &F1, P1);
  KeyFrame* K2 = build(k2, Tcw2, &F2, P2);
  restore_statics(cam);
  cv::Mat F(3, 3, CV_32F);
  memcpy(F.data, F12, 36);
  std::vector<std::pair<size_t, size_t>> pairs;
  ORBmatcher m(nnratio, check_ori != 0);
  *nmatches_out = m.SearchForTriangulation(K1, K2, F, pairs, only_stereo != 0);
  for (int i = 0; i < k1->n; ++i) matches12[i] = -1;
  for (auto& pr : pairs) matches12[pr.first] = (int32_t)pr.second;
  if (epipole_out) {   // the same expressions as :835-839
    cv::Mat Cw = K1->GetCameraCenter();
    cv::Mat R2w = K2->GetRotation();
    cv::Mat t2w = K2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    epipole_out[0] = K2->fx * C2.at<float>(0) * invz + K2->cx;
    epipole_out[1] = K2->fy * C2.at<float>(1) * invz + K2->cy;
  }
  delete K1; delete K2; delete F1; delete F2;
  return 0;
}

// KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:659-698) + ORBmatcher::DescriptorDistance for a list of queries: the
// candidate walk the BEST searches share, without any gate.  best_dist = INT_MAX when no candidate.
int X(kf_best)(const OrbmFrame* kf, const OrbmQueries* q, int32_t* best_idx, int32_t* best_dist) {
  std::lock_guard<std::mutex> lk(g_mu);
  Frame* FK;
  KeyFrame* K = make_keyframe(kf, nullptr, &FK);
  restore_statics(kf);
  for (int i = 0; i < q->n; ++i) {
    best_idx[i] = -1; best_dist[i] = 2147483647;
    if (!q->valid[i]) continue;
    const std::vector<size_t> vIndices = K->GetFeaturesInArea(q->u[i], q->v[i], q->radius[i]);
    const cv::Mat dMP = desc_row(q->desc + 32 * (size_t)i);
    for (size_t idx : vIndices) {
      const int lvl = K->mvKeysUn[idx].octave;
      if (lvl < q->min_level[i] || lvl > q->max_level[i]) continue;
      const int dist = ORBmatcher::DescriptorDistance(dMP, K->mDescriptors.row((int)idx));
      if (dist < best_dist[i]) { best_dist[i] = dist; best_idx[i] = (int32_t)idx; }
    }
  }
  delete K; delete FK;
  return 0;
}


// Frame::ComputeBoW (src/Frame.cc:546-555) on a real Frame with a vocabulary built from flat arrays (node 0 = root,
// parent[i] < i, leaves = words in ascending id): -> the BowVector as (word, value) pairs in map order and the
// FeatureVector flattened (node ids ascending, feature indices per node in insertion order).
int X(bow_transform)(int k, int L, int n_nodes, const int32_t* parent, const uint8_t* node_desc, const double* weight,
                     const uint8_t* desc, int n, int levelsup_unused, uint32_t* words, double* values, int* n_words,
                     uint32_t* node_ids, int32_t* node_off, uint32_t* idx, int* n_nodes_out) {
  (void)levelsup_unused;   // ComputeBoW hard-codes 4
  std::lock_guard<std::mutex> lk(g_mu);
  ORBVocabulary voc(k, L, DBoW2::TF_IDF, DBoW2::L1_NORM);
  std::vector<DBoW2::NodeId> par(n_nodes);
  std::vector<cv::Mat> nd(n_nodes);
  std::vector<DBoW2::WordValue> w(weight, weight + n_nodes);
  for (int i = 0; i < n_nodes; ++i) { par[i] = (DBoW2::NodeId)parent[i]; nd[i] = desc_row(node_desc + 32 * (size_t)i); }
  voc.build(k, L, par, nd, w);
  std::vector<float> zf(n, 0.f);
  std::vector<int32_t> zi(n, 0);
  OrbmBow b;
  memset(&b, 0, sizeof(b));
  b.n = n; b.desc = desc; b.angle = zf.data();
  OrbmFrame v = bow_as_frame(&b, zf, zi);
  Frame* F = make_frame(&v);
  F->mpORBvocabulary = &voc;
  F->ComputeBoW();
  int nw = 0;
  for (auto& kv : F->mBowVec) { words[nw] = kv.first; values[nw] = kv.second; ++nw; }
  *n_words = nw;
  int nn = 0, ni = 0;
  node_off[0] = 0;
  for (auto& kv : F->mFeatVec) {
    node_ids[nn] = kv.first;
    for (unsigned int f : kv.second) idx[ni++] = f;
    node_off[++nn] = ni;
  }
  *n_nodes_out = nn;
  F->mpORBvocabulary = nullptr;
  delete F;
  return 0;
}

}  // extern "C"
