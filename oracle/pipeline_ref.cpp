// oracle/pipeline_ref.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// CPU baseline driver: runs the oracle's extract -> ComputeStereoFromRGBD/UnprojectStereo -> SearchByProjection
// over a batch of frames on `nthreads` host threads (frame-parallel, one extractor instance per thread, which is
// how the reference itself parallelises extraction for stereo, src/Frame.cc:121-124) and times it with
// std::chrono::steady_clock like the reference's own driver does (perfect/Examples/RGB-D/rgbd_tum.cc:92-111).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/b200orb.h"

extern "C" {
void* orb_ref_create(int, float, int, int, int);
void orb_ref_destroy(void*);
int orb_ref_extract(void*, const uint8_t*, int, int, int, void*, uint8_t*, int, int*);
void orb_ref_tables(void*, float*, float*, float*, float*, int*, int*);
int match_ref_projection_last(const OrbmFrame*, const OrbmLast*, float, int, float, int, int32_t*, int*);
void frame_ref_stereo_unproject(const float*, int, int, const float*, int, int, const float*, float, float, float,
                                float, float, float*, float*, float*, uint8_t*);
void occ_ref_default_params(OcmParams*);
void* occ_ref_create(const OcmParams*);
void occ_ref_destroy(void*);
int occ_ref_insert_keyframe(void*, const float*, const uint8_t*, int, int, const float*, float, float, float, float,
                            const uint8_t*);
long long occ_ref_num_leaves(void*);

// Returns elapsed seconds; nkp[n], nmatch[n] receive per-frame counts (nmatch[0] = 0).
double pipeline_ref_run(const uint8_t* gray, const float* depth, const float* Tcw, int n, int rows, int cols,
                        int nfeatures, float scale, int nlevels, int ini_th, int min_th, float fx, float fy, float cx,
                        float cy, float bf, float th, float nnratio, int check_ori, int last_obs, int nthreads,
                        int* nkp, int* nmatch, const uint8_t* rgb, int kf_every, long long* leaves_out) {
  const int cap = nfeatures + 3 * nlevels + 64;
  std::vector<OrbxKeyPoint> kps((size_t)n * cap);
  std::vector<uint8_t> desc((size_t)n * cap * 32);
  std::vector<float> sf(nlevels), tmp(nlevels);
  std::vector<int> itmp(nlevels), um(16);
  {
    void* e = orb_ref_create(nfeatures, scale, nlevels, ini_th, min_th);
    orb_ref_tables(e, sf.data(), tmp.data(), tmp.data(), tmp.data(), itmp.data(), um.data());
    orb_ref_destroy(e);
  }
  const size_t px = (size_t)rows * cols;
  auto t0 = std::chrono::steady_clock::now();
  // the dense-mapping thread of the reference (src/pointcloudmapping.cc:43: its own std::thread) runs beside tracking
  std::thread mapper;
  if (rgb && kf_every > 0) {
    mapper = std::thread([&]() {
      OcmParams op;
      occ_ref_default_params(&op);
      void* m = occ_ref_create(&op);
      for (int i = 0; i < n; i += kf_every)
        occ_ref_insert_keyframe(m, depth + px * i, rgb + px * 3 * i, rows, cols, Tcw + 16 * i, fx, fy, cx, cy, nullptr);
      if (leaves_out) *leaves_out = occ_ref_num_leaves(m);
      occ_ref_destroy(m);
    });
  }
  {
    std::atomic<int> next(0);
    std::vector<std::thread> th_;
    for (int t = 0; t < nthreads; ++t)
      th_.emplace_back([&]() {
        void* e = orb_ref_create(nfeatures, scale, nlevels, ini_th, min_th);
        for (int i = next++; i < n; i = next++)
          nkp[i] = orb_ref_extract(e, gray + px * i, rows, cols, cols, &kps[(size_t)i * cap], &desc[(size_t)i * cap * 32],
                                   cap, nullptr);
        orb_ref_destroy(e);
      });
    for (auto& t : th_) t.join();
  }
  nmatch[0] = 0;
  {
    std::atomic<int> next(1);
    std::vector<std::thread> th_;
    for (int t = 0; t < nthreads; ++t)
      th_.emplace_back([&]() {
        std::vector<float> cxv(cap), cyv(cap), cang(cap), cur(cap), cdep(cap), cxw(cap * 3), lur(cap), ldep(cap),
            lxw(cap * 3), lang(cap);
        std::vector<int32_t> coct(cap), loct(cap), lobs(cap, last_obs), out(cap);
        std::vector<uint8_t> cval(cap), lval(cap);
        for (int i = next++; i < n; i = next++) {
          const int nc = nkp[i], nl = nkp[i - 1];
          const OrbxKeyPoint* kc = &kps[(size_t)i * cap];
          const OrbxKeyPoint* kl = &kps[(size_t)(i - 1) * cap];
          frame_ref_stereo_unproject((const float*)kc, 7, nc, depth + px * i, rows, cols, Tcw + 16 * i, fx, fy, cx, cy,
                                     bf, cur.data(), cdep.data(), cxw.data(), cval.data());
          frame_ref_stereo_unproject((const float*)kl, 7, nl, depth + px * (i - 1), rows, cols, Tcw + 16 * (i - 1), fx,
                                     fy, cx, cy, bf, lur.data(), ldep.data(), lxw.data(), lval.data());
          for (int k = 0; k < nc; ++k) { cxv[k] = kc[k].x; cyv[k] = kc[k].y; cang[k] = kc[k].angle; coct[k] = kc[k].octave; }
          for (int k = 0; k < nl; ++k) { lang[k] = kl[k].angle; loct[k] = kl[k].octave; }
          OrbmFrame F;
          memset(&F, 0, sizeof(F));
          F.n = nc; F.x = cxv.data(); F.y = cyv.data(); F.octave = coct.data(); F.angle = cang.data();
          F.uright = cur.data(); F.desc = &desc[(size_t)i * cap * 32]; F.mp_obs = nullptr;
          memcpy(F.Tcw, Tcw + 16 * i, 64);
          F.fx = fx; F.fy = fy; F.cx = cx; F.cy = cy; F.bf = bf; F.b = bf / fx;
          F.min_x = 0; F.max_x = (float)cols; F.min_y = 0; F.max_y = (float)rows;
          F.scale_factors = sf.data(); F.nlevels = nlevels;
          OrbmLast L;
          memset(&L, 0, sizeof(L));
          L.n = nl; L.xw = lxw.data(); L.valid = lval.data(); L.octave = loct.data(); L.angle = lang.data();
          L.mp_desc = &desc[(size_t)(i - 1) * cap * 32]; L.mp_obs = lobs.data();
          memcpy(L.Tcw, Tcw + 16 * (i - 1), 64);
          match_ref_projection_last(&F, &L, th, 0, nnratio, check_ori, out.data(), &nmatch[i]);
        }
      });
    for (auto& t : th_) t.join();
  }
  if (mapper.joinable()) mapper.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
}
