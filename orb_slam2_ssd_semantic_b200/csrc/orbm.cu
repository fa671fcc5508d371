// orbm.cu -- host side + C-ABI of the B200 ORB matcher (reference: src/ORBmatcher.cc, include/ORBmatcher.h).
#include <algorithm>
#include <cmath>
#include <new>
#include <vector>

#include "orbm_host.h"
#include "orbm_extra.cuh"

using namespace b200;

orbm::~orbm() {
  DeviceGuard g(device);
  if (d_arena) cudaFree(d_arena);
  if (stream) cudaStreamDestroy(stream);
}

int orbm::reserve(size_t bytes) {
  if (bytes <= arena_bytes) return B200ORB_OK;
  if (d_arena) { cudaStreamSynchronize(stream); cudaFree(d_arena); d_arena = nullptr; arena_bytes = 0; }
  const size_t want = align_up_sz(bytes + bytes / 2, 1 << 20);
  B200_CUDA(cudaMalloc(&d_arena, want));
  arena_bytes = want;
  return B200ORB_OK;
}

namespace {
struct Carver {   // sequential sub-allocation out of the arena, 256-byte aligned
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  template <class T>
  T* take(size_t n) {
    T* p = (T*)(base + off);
    off = align_up_sz(off + std::max<size_t>(n, 1) * sizeof(T), 256);
    return p;
  }
};

int resolve_smem(int cmax, size_t* smem, const void* func) {
  *smem = (size_t)cmax * 5 + 16;
  if (*smem > 200 * 1024) { set_error("too many keypoints per frame for the matcher's shared memory"); return B200ORB_EINVAL; }
  if (*smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
  return B200ORB_OK;
}
}  // namespace

int launch_match_last_unfused(const CurView& cv_in, const LastView& lv, const MatchCam& cam, int npairs, int* d_goff,
                              int* d_gidx, unsigned* d_list, int* d_count, int* d_accepted, int* d_cur2last,
                              int* d_nmatch, int cmax, int lmax, cudaStream_t stream, long long* launches);

// Enqueue the LAST search of `npairs` (cur, last) instances.
int launch_match_last(const CurView& cv_in, const LastView& lv, const MatchCam& cam, int npairs, int* d_goff, int* d_gidx,
                      unsigned* d_list, int* d_count, int* d_accepted, int* d_cur2last, int* d_nmatch, int cmax,
                      int lmax, cudaStream_t stream, long long* launches) {
  const int lmax16 = align_up(lmax, 16);
  const size_t smem = mf_smem_bytes(cmax, lmax16);
  if (smem > 200 * 1024 || cmax > 4096)   // frame too large to stage in shared memory (or for the packed resolve key):
                                          // separate grid / candidate / resolve kernels
    return launch_match_last_unfused(cv_in, lv, cam, npairs, d_goff, d_gidx, d_list, d_count, d_accepted, d_cur2last,
                                     d_nmatch, cmax, lmax, stream, launches);
  B200_CUDA(cudaFuncSetAttribute(k_match_last_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ListView lsv{d_list, d_count};
  k_match_last_fused<<<npairs, MF_THREADS, smem, stream>>>(cv_in, lv, cam, lsv, d_accepted, d_cur2last, d_nmatch, cmax, lmax16);
  if (launches) *launches += 1;
  B200_CUDA(cudaGetLastError());
  return B200ORB_OK;
}

// The unfused three-kernel path (grid / candidates / resolve as separate launches); kept for A/B timing.
int launch_match_last_unfused(const CurView& cv_in, const LastView& lv, const MatchCam& cam, int npairs, int* d_goff,
                              int* d_gidx, unsigned* d_list, int* d_count, int* d_accepted, int* d_cur2last,
                              int* d_nmatch, int cmax, int lmax, cudaStream_t stream, long long* launches) {
  CurView cv = cv_in;
  k_grid_build<<<npairs, 256, 0, stream>>>(cv.x, cv.y, cv.n, cv.stride, cam.min_x, cam.max_x, cam.min_y, cam.max_y, d_goff,
                                          d_gidx);
  cv.goff = d_goff; cv.gidx = d_gidx;
  ListView lsv{d_list, d_count};
  k_cand_last<<<dim3((lmax + CAND_WARPS - 1) / CAND_WARPS, npairs), CAND_WARPS * 32, 0, stream>>>(cv, lv, cam, lsv);
  size_t smem;
  B200_CHECK(resolve_smem(cmax, &smem, (const void*)k_resolve_last));
  k_resolve_last<<<npairs, 32, smem, stream>>>(cv, lv, cam, lsv, d_accepted, d_cur2last, d_nmatch, cmax);
  if (launches) *launches += 3;
  B200_CUDA(cudaGetLastError());
  return B200ORB_OK;
}

extern "C" {

#ifdef B200ORB_TIMING
int orbm_debug_read(unsigned long long* out16) {   // profiling builds only: phase cycle counters of k_match_last_fused
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, g_mf_dbg, sizeof(unsigned long long) * 16) != cudaSuccess) return B200ORB_ECUDA;
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_mf_dbg, z, sizeof(z));
  return B200ORB_OK;
}
#endif

int orbm_hamming(const uint8_t a[32], const uint8_t b[32]) {   // DescriptorDistance, src/ORBmatcher.cc:1968-1984
  int d = 0;
  for (int i = 0; i < 32; i += 4) {
    uint32_t x, y;
    memcpy(&x, a + i, 4);
    memcpy(&y, b + i, 4);
    d += __builtin_popcount(x ^ y);
  }
  return d;
}

int orbm_create(int device, orbm_t** out) {
  if (!out) { set_error("null argument"); return B200ORB_EINVAL; }
  *out = nullptr;
  B200_CHECK(check_device(device));
  DeviceGuard g(device);
  orbm* h = new (std::nothrow) orbm();
  if (!h) { set_error("out of host memory"); return B200ORB_EINVAL; }
  h->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); delete h; return B200ORB_ECUDA; }
  *out = h;
  return B200ORB_OK;
}
void orbm_destroy(orbm_t* h) { delete h; }
long long orbm_launch_count(const orbm_t* h) { return h ? h->launches : 0; }

static int fill_cam(const OrbmFrame* f, float th, int mono, float nnratio, int check_ori, MatchCam* cam) {
  if (f->nlevels < 1 || f->nlevels > MAX_LEVELS || !f->scale_factors) { set_error("bad scale factor table"); return B200ORB_EINVAL; }
  if (!(f->max_x > f->min_x) || !(f->max_y > f->min_y)) { set_error("bad image bounds"); return B200ORB_EINVAL; }
  memset(cam, 0, sizeof(*cam));
  cam->fx = f->fx; cam->fy = f->fy; cam->cx = f->cx; cam->cy = f->cy; cam->bf = f->bf; cam->b = f->b;
  cam->min_x = f->min_x; cam->max_x = f->max_x; cam->min_y = f->min_y; cam->max_y = f->max_y;
  for (int i = 0; i < f->nlevels; ++i) cam->sf[i] = f->scale_factors[i];
  cam->th = th; cam->nnratio = nnratio; cam->mono = mono; cam->check_ori = check_ori; cam->nlevels = f->nlevels;
  return B200ORB_OK;
}

#define UP(dst, src, count, T)                                                                              \
  B200_CUDA(cudaMemcpyAsync((dst), (src), sizeof(T) * (size_t)(count), cudaMemcpyHostToDevice, h->stream))

// uploads the current-frame arrays; fills cv (single instance)
static int upload_cur(orbm* h, Carver& cv_mem, const OrbmFrame* cur, CurView* cv, int** d_goff, int** d_gidx) {
  const size_t nc = cur->n;
  float* d_cx = cv_mem.take<float>(nc); float* d_cy = cv_mem.take<float>(nc); float* d_cang = cv_mem.take<float>(nc);
  float* d_cur = cv_mem.take<float>(nc); int* d_coct = cv_mem.take<int>(nc); uint8_t* d_cdesc = cv_mem.take<uint8_t>(nc * 32);
  int* d_cobs = cur->mp_obs ? cv_mem.take<int>(nc) : nullptr;
  int* d_cn = cv_mem.take<int>(1); float* d_cT = cv_mem.take<float>(16);
  *d_goff = cv_mem.take<int>(GRID_CELLS + 1); *d_gidx = cv_mem.take<int>(nc);
  UP(d_cx, cur->x, nc, float); UP(d_cy, cur->y, nc, float); UP(d_cang, cur->angle, nc, float);
  UP(d_cur, cur->uright, nc, float); UP(d_coct, cur->octave, nc, int); UP(d_cdesc, cur->desc, nc * 32, uint8_t);
  if (d_cobs) UP(d_cobs, cur->mp_obs, nc, int);
  const int cn = cur->n;
  UP(d_cn, &cn, 1, int); UP(d_cT, cur->Tcw, 16, float);
  memset(cv, 0, sizeof(*cv));
  cv->x = d_cx; cv->y = d_cy; cv->ang = d_cang; cv->uright = d_cur; cv->oct = d_coct; cv->desc = d_cdesc; cv->obs = d_cobs;
  cv->n = d_cn; cv->Tcw = d_cT; cv->stride = nc; cv->goff = *d_goff; cv->gidx = *d_gidx;
  return B200ORB_OK;
}

static int check_cur(const OrbmFrame* cur) {
  if (cur->n < 0 || cur->n >= (1 << 20)) { set_error("bad keypoint count"); return B200ORB_EINVAL; }
  if (cur->n > 0 && (!cur->x || !cur->y || !cur->octave || !cur->angle || !cur->uright || !cur->desc)) {
    set_error("null keypoint array");
    return B200ORB_EINVAL;
  }
  return B200ORB_OK;
}

int orbm_search_by_projection_last(orbm_t* h, const OrbmFrame* cur, const OrbmLast* last, float th, int mono,
                                   float nnratio, int check_ori, int32_t* cur2last, int* nmatches) {
  if (!h || !cur || !last || !cur2last || !nmatches) { set_error("null argument"); return B200ORB_EINVAL; }
  B200_CHECK(check_cur(cur));
  if (last->n < 0) { set_error("bad keypoint count"); return B200ORB_EINVAL; }
  *nmatches = 0;
  if (cur->n == 0) return B200ORB_OK;
  if (last->n == 0) {
    for (int j = 0; j < cur->n; ++j) cur2last[j] = (cur->mp_obs && cur->mp_obs[j] >= 0) ? -2 : -1;
    return B200ORB_OK;
  }
  for (int i = 0; i < last->n; ++i)
    if (last->valid[i] && (last->octave[i] < 0 || last->octave[i] >= cur->nlevels)) { set_error("octave out of range"); return B200ORB_EINVAL; }
  MatchCam cam;
  B200_CHECK(fill_cam(cur, th, mono, nnratio, check_ori, &cam));
  cam.last_obs_default = 0;
  DeviceGuard g(h->device);
  const size_t nc = cur->n, nl = last->n;
  const size_t need = nc * (4 * 4 + 4 + 32 + 4 + 4 + 4) + nl * (12 + 1 + 4 + 4 + 32 + 4 + 4 * LCAP + 4 + 4) + 4 * (GRID_CELLS + 1) +
                      64 * 256 + 4096;
  B200_CHECK(h->reserve(need));
  Carver cm(h->d_arena);
  CurView cv;
  int *d_goff, *d_gidx;
  B200_CHECK(upload_cur(h, cm, cur, &cv, &d_goff, &d_gidx));
  float* d_lxw = cm.take<float>(nl * 3); uint8_t* d_lvalid = cm.take<uint8_t>(nl); int* d_loct = cm.take<int>(nl);
  float* d_lang = cm.take<float>(nl); uint8_t* d_ldesc = cm.take<uint8_t>(nl * 32);
  int* d_lobs = last->mp_obs ? cm.take<int>(nl) : nullptr;
  int* d_ln = cm.take<int>(1); float* d_lT = cm.take<float>(16);
  int* d_out = cm.take<int>(nc); int* d_nm = cm.take<int>(1);
  unsigned* d_list = cm.take<unsigned>(nl * LCAP); int* d_count = cm.take<int>(nl); int* d_acc = cm.take<int>(nl);
  UP(d_lxw, last->xw, nl * 3, float); UP(d_lvalid, last->valid, nl, uint8_t); UP(d_loct, last->octave, nl, int);
  UP(d_lang, last->angle, nl, float); UP(d_ldesc, last->mp_desc, nl * 32, uint8_t);
  if (d_lobs) UP(d_lobs, last->mp_obs, nl, int);
  const int ln = last->n;
  UP(d_ln, &ln, 1, int); UP(d_lT, last->Tcw, 16, float);
  LastView lv{d_lxw, d_lvalid, d_loct, d_lang, d_ldesc, d_lobs, d_ln, d_lT, nl};
  B200_CHECK(launch_match_last(cv, lv, cam, 1, d_goff, d_gidx, d_list, d_count, d_acc, d_out, d_nm, align_up((int)nc, 16),
                               (int)nl, h->stream, &h->launches));
  B200_CUDA(cudaMemcpyAsync(cur2last, d_out, sizeof(int) * nc, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(nmatches, d_nm, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int orbm_search_by_projection_points(orbm_t* h, const OrbmFrame* f, const OrbmTrackPoints* pts, float th, float nnratio,
                                     int32_t* f2pt, int* nmatches) {
  if (!h || !f || !pts || !f2pt || !nmatches) { set_error("null argument"); return B200ORB_EINVAL; }
  B200_CHECK(check_cur(f));
  if (pts->n < 0) { set_error("bad point count"); return B200ORB_EINVAL; }
  *nmatches = 0;
  if (f->n == 0) return B200ORB_OK;
  if (pts->n == 0) {
    for (int j = 0; j < f->n; ++j) f2pt[j] = (f->mp_obs && f->mp_obs[j] >= 0) ? -2 : -1;
    return B200ORB_OK;
  }
  for (int i = 0; i < pts->n; ++i)
    if (pts->track_in_view[i] && (pts->scale_level[i] < 0 || pts->scale_level[i] >= f->nlevels)) {
      set_error("scale level out of range");
      return B200ORB_EINVAL;
    }
  MatchCam cam;
  B200_CHECK(fill_cam(f, th, 0, nnratio, 0, &cam));
  cam.last_obs_default = 1;
  DeviceGuard g(h->device);
  const size_t nc = f->n, np = pts->n;
  const size_t need = nc * (4 * 4 + 4 + 32 + 4 + 4 + 4) + np * (1 + 16 + 4 + 32 + 4 + 4 * LCAP + 4) + 4 * (GRID_CELLS + 1) + 64 * 256 + 4096;
  B200_CHECK(h->reserve(need));
  Carver cm(h->d_arena);
  CurView cv;
  int *d_goff, *d_gidx;
  B200_CHECK(upload_cur(h, cm, f, &cv, &d_goff, &d_gidx));
  uint8_t* d_inv = cm.take<uint8_t>(np); float* d_px = cm.take<float>(np); float* d_py = cm.take<float>(np);
  float* d_pxr = cm.take<float>(np); float* d_vc = cm.take<float>(np); int* d_lvl = cm.take<int>(np);
  uint8_t* d_desc = cm.take<uint8_t>(np * 32); int* d_obs = pts->mp_obs ? cm.take<int>(np) : nullptr;
  unsigned* d_list = cm.take<unsigned>(np * LCAP); int* d_count = cm.take<int>(np);
  int* d_out = cm.take<int>(nc); int* d_nm = cm.take<int>(1);
  UP(d_inv, pts->track_in_view, np, uint8_t); UP(d_px, pts->proj_x, np, float); UP(d_py, pts->proj_y, np, float);
  UP(d_pxr, pts->proj_xr, np, float); UP(d_vc, pts->view_cos, np, float); UP(d_lvl, pts->scale_level, np, int);
  UP(d_desc, pts->mp_desc, np * 32, uint8_t);
  if (d_obs) UP(d_obs, pts->mp_obs, np, int);
  PointsView pv{d_inv, d_px, d_py, d_pxr, d_vc, d_lvl, d_desc, d_obs, (int)np};
  ListView lsv{d_list, d_count};
  k_grid_build<<<1, 256, 0, h->stream>>>(cv.x, cv.y, cv.n, cv.stride, cam.min_x, cam.max_x, cam.min_y, cam.max_y, d_goff, d_gidx);
  k_cand_points<<<((int)np + CAND_WARPS - 1) / CAND_WARPS, CAND_WARPS * 32, 0, h->stream>>>(cv, pv, cam, lsv);
  size_t smem;
  const int cmax = align_up((int)nc, 16);
  B200_CHECK(resolve_smem(cmax, &smem, (const void*)k_resolve_points));
  k_resolve_points<<<1, 32, smem, h->stream>>>(cv, pv, cam, lsv, d_out, d_nm, cmax);
  h->launches += 3;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(f2pt, d_out, sizeof(int) * nc, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(nmatches, d_nm, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int orbm_search_projected(orbm_t* h, const OrbmFrame* cur, const OrbmQueries* q, int max_dist, int claim_rule,
                          int check_ori, int32_t* cur2q, int* nmatches) {
  if (!h || !cur || !q || !cur2q || !nmatches) { set_error("null argument"); return B200ORB_EINVAL; }
  B200_CHECK(check_cur(cur));
  if (q->n < 0) { set_error("bad query count"); return B200ORB_EINVAL; }
  *nmatches = 0;
  if (cur->n == 0) return B200ORB_OK;
  if (q->n == 0) {
    for (int j = 0; j < cur->n; ++j) cur2q[j] = (cur->mp_obs && cur->mp_obs[j] >= 0) ? -2 : -1;
    return B200ORB_OK;
  }
  if (!q->valid || !q->u || !q->v || !q->radius || !q->min_level || !q->max_level || !q->desc || (check_ori && !q->angle)) {
    set_error("null query array");
    return B200ORB_EINVAL;
  }
  MatchCam cam;
  B200_CHECK(fill_cam(cur, 0.f, 0, 0.f, check_ori, &cam));
  DeviceGuard g(h->device);
  const size_t nc = cur->n, nq = q->n;
  const size_t need = nc * (4 * 4 + 4 + 32 + 4 + 4 + 4) + nq * (1 + 4 * 5 + 8 + 32 + 4 + 4 * LCAP + 8) + 4 * (GRID_CELLS + 1) + 64 * 256 + 4096;
  B200_CHECK(h->reserve(need));
  Carver cm(h->d_arena);
  CurView cv;
  int *d_goff, *d_gidx;
  B200_CHECK(upload_cur(h, cm, cur, &cv, &d_goff, &d_gidx));
  uint8_t* d_valid = cm.take<uint8_t>(nq); float* d_u = cm.take<float>(nq); float* d_v = cm.take<float>(nq);
  float* d_r = cm.take<float>(nq); int* d_mn = cm.take<int>(nq); int* d_mx = cm.take<int>(nq);
  float* d_ur = q->uright ? cm.take<float>(nq) : nullptr; float* d_ang = cm.take<float>(nq);
  int* d_obs = q->obs ? cm.take<int>(nq) : nullptr; uint8_t* d_desc = cm.take<uint8_t>(nq * 32);
  unsigned* d_list = cm.take<unsigned>(nq * LCAP); int* d_count = cm.take<int>(nq); int* d_acc = cm.take<int>(nq);
  int* d_out = cm.take<int>(nc); int* d_nm = cm.take<int>(1);
  UP(d_valid, q->valid, nq, uint8_t); UP(d_u, q->u, nq, float); UP(d_v, q->v, nq, float); UP(d_r, q->radius, nq, float);
  UP(d_mn, q->min_level, nq, int); UP(d_mx, q->max_level, nq, int); UP(d_desc, q->desc, nq * 32, uint8_t);
  if (d_ur) UP(d_ur, q->uright, nq, float);
  if (q->angle) UP(d_ang, q->angle, nq, float);
  if (d_obs) UP(d_obs, q->obs, nq, int);
  QueriesView qv{d_valid, d_u, d_v, d_r, d_ur, d_ang, d_mn, d_mx, d_obs, d_desc, (int)nq};
  ListView lsv{d_list, d_count};
  k_grid_build<<<1, 256, 0, h->stream>>>(cv.x, cv.y, cv.n, cv.stride, cam.min_x, cam.max_x, cam.min_y, cam.max_y, d_goff, d_gidx);
  k_cand_generic<<<((int)nq + CAND_WARPS - 1) / CAND_WARPS, CAND_WARPS * 32, 0, h->stream>>>(cv, qv, cam, claim_rule, lsv);
  size_t smem;
  const int cmax = align_up((int)nc, 16);
  B200_CHECK(resolve_smem(cmax, &smem, (const void*)k_resolve_generic));
  k_resolve_generic<<<1, 32, smem, h->stream>>>(cv, qv, cam, max_dist, claim_rule, lsv, d_acc, d_out, d_nm, cmax);
  h->launches += 3;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(cur2q, d_out, sizeof(int) * nc, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(nmatches, d_nm, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int orbm_search_best(orbm_t* h, const OrbmFrame* kf, const OrbmQueries* q, int gate, const float* inv_level_sigma2,
                     int32_t* best_idx, int32_t* best_dist) {
  if (!h || !kf || !q || !best_idx || !best_dist) { set_error("null argument"); return B200ORB_EINVAL; }
  B200_CHECK(check_cur(kf));
  if (q->n < 0) { set_error("bad query count"); return B200ORB_EINVAL; }
  for (int i = 0; i < q->n; ++i) { best_idx[i] = -1; best_dist[i] = 0x7fffffff; }
  if (kf->n == 0 || q->n == 0) return B200ORB_OK;
  if (!q->valid || !q->u || !q->v || !q->radius || !q->min_level || !q->max_level || !q->desc) { set_error("null query array"); return B200ORB_EINVAL; }
  if (gate == 1 && (!inv_level_sigma2 || !q->uright)) { set_error("the Fuse gate needs inv_level_sigma2 and the predicted right coordinates"); return B200ORB_EINVAL; }
  if (gate != 0 && gate != 1) { set_error("bad gate"); return B200ORB_EINVAL; }
  MatchCam cam;
  B200_CHECK(fill_cam(kf, 0.f, 0, 0.f, 0, &cam));
  DeviceGuard g(h->device);
  const size_t nc = kf->n, nq = q->n;
  const size_t need = nc * (4 * 4 + 4 + 32 + 4 + 4 + 4) + nq * (1 + 4 * 6 + 8 + 32 + 8) + 4 * (GRID_CELLS + 1) + 64 * 256 + 8192;
  B200_CHECK(h->reserve(need));
  Carver cm(h->d_arena);
  CurView cv;
  int *d_goff, *d_gidx;
  OrbmFrame kf2 = *kf;
  kf2.mp_obs = nullptr;
  B200_CHECK(upload_cur(h, cm, &kf2, &cv, &d_goff, &d_gidx));
  uint8_t* d_valid = cm.take<uint8_t>(nq); float* d_u = cm.take<float>(nq); float* d_v = cm.take<float>(nq);
  float* d_r = cm.take<float>(nq); int* d_mn = cm.take<int>(nq); int* d_mx = cm.take<int>(nq);
  float* d_ur = q->uright ? cm.take<float>(nq) : nullptr; uint8_t* d_desc = cm.take<uint8_t>(nq * 32);
  float* d_is2 = cm.take<float>(MAX_LEVELS);
  int* d_bi = cm.take<int>(nq); int* d_bd = cm.take<int>(nq);
  UP(d_valid, q->valid, nq, uint8_t); UP(d_u, q->u, nq, float); UP(d_v, q->v, nq, float); UP(d_r, q->radius, nq, float);
  UP(d_mn, q->min_level, nq, int); UP(d_mx, q->max_level, nq, int); UP(d_desc, q->desc, nq * 32, uint8_t);
  if (d_ur) UP(d_ur, q->uright, nq, float);
  if (gate == 1) UP(d_is2, inv_level_sigma2, kf->nlevels, float);
  QueriesView qv{d_valid, d_u, d_v, d_r, d_ur, nullptr, d_mn, d_mx, nullptr, d_desc, (int)nq};
  BestGate bg{gate, d_is2, cv.uright};
  k_grid_build<<<1, 256, 0, h->stream>>>(cv.x, cv.y, cv.n, cv.stride, cam.min_x, cam.max_x, cam.min_y, cam.max_y, d_goff, d_gidx);
  k_best_generic<<<((int)nq + CAND_WARPS - 1) / CAND_WARPS, CAND_WARPS * 32, 0, h->stream>>>(cv, qv, cam, bg, d_bi, d_bd);
  h->launches += 2;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(best_idx, d_bi, sizeof(int) * nq, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(best_dist, d_bd, sizeof(int) * nq, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int orbm_search_for_initialization(orbm_t* h, const OrbmFrame* f1, const OrbmFrame* f2, float* prev_xy, int window,
                                   float nnratio, int check_ori, int32_t* matches12, int* nmatches) {
  if (!h || !f1 || !f2 || !prev_xy || !matches12 || !nmatches) { set_error("null argument"); return B200ORB_EINVAL; }
  B200_CHECK(check_cur(f1));
  B200_CHECK(check_cur(f2));
  *nmatches = 0;
  for (int i = 0; i < f1->n; ++i) matches12[i] = -1;
  if (f1->n == 0 || f2->n == 0) return B200ORB_OK;
  MatchCam cam;
  B200_CHECK(fill_cam(f2, 0.f, 0, nnratio, check_ori, &cam));
  DeviceGuard g(h->device);
  const size_t n1 = f1->n, n2 = f2->n;
  const size_t need = n2 * (4 * 4 + 4 + 32 + 4 + 4 + 4 + 8) + n1 * (4 + 4 + 32 + 8 + 4 + 4) + 4 * (GRID_CELLS + 1) + 64 * 256 + 8192;
  B200_CHECK(h->reserve(need));
  Carver cm(h->d_arena);
  CurView cv;
  int *d_goff, *d_gidx;
  OrbmFrame f2c = *f2;
  f2c.mp_obs = nullptr;
  B200_CHECK(upload_cur(h, cm, &f2c, &cv, &d_goff, &d_gidx));
  int* d_oct1 = cm.take<int>(n1); float* d_ang1 = cm.take<float>(n1); uint8_t* d_desc1 = cm.take<uint8_t>(n1 * 32);
  float* d_prev = cm.take<float>(n1 * 2); int* d_md = cm.take<int>(n2); int* d_m21 = cm.take<int>(n2);
  int* d_bin = cm.take<int>(n1); int* d_out = cm.take<int>(n1); int* d_nm = cm.take<int>(1);
  UP(d_oct1, f1->octave, n1, int); UP(d_ang1, f1->angle, n1, float); UP(d_desc1, f1->desc, n1 * 32, uint8_t);
  UP(d_prev, prev_xy, n1 * 2, float);
  InitView iv{d_oct1, d_ang1, d_desc1, d_prev, (int)n1, d_md, d_m21, d_bin};
  k_grid_build<<<1, 256, 0, h->stream>>>(cv.x, cv.y, cv.n, cv.stride, cam.min_x, cam.max_x, cam.min_y, cam.max_y, d_goff, d_gidx);
  k_init_search<<<1, 32, 0, h->stream>>>(cv, iv, cam, window, nnratio, check_ori, d_out, d_nm);
  h->launches += 2;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(matches12, d_out, sizeof(int) * n1, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(prev_xy, d_prev, sizeof(float) * 2 * n1, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(nmatches, d_nm, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int orbm_search_for_triangulation(orbm_t* h, const OrbmTriKF* k1, const OrbmTriKF* k2, const float F12[9], float ex, float ey,
                                  const float* scale_factors2, const float* level_sigma2_2, int nlevels, int only_stereo,
                                  int check_ori, int32_t* matches12, int* nmatches) {
  if (!h || !k1 || !k2 || !F12 || !scale_factors2 || !level_sigma2_2 || !matches12 || !nmatches) { set_error("null argument"); return B200ORB_EINVAL; }
  if (k1->n < 0 || k2->n < 0 || k1->n_nodes < 0 || k2->n_nodes < 0 || k2->n >= (1 << 22) || nlevels < 1 || nlevels > MAX_LEVELS) { set_error("bad counts"); return B200ORB_EINVAL; }
  *nmatches = 0;
  for (int i = 0; i < k1->n; ++i) matches12[i] = -1;
  if (k1->n == 0 || k2->n == 0 || k1->n_nodes == 0 || k2->n_nodes == 0) return B200ORB_OK;
  for (int i = 0; i < k2->n; ++i)
    if (k2->octave[i] < 0 || k2->octave[i] >= nlevels) { set_error("octave out of range"); return B200ORB_EINVAL; }
  // merge-join of the FeatureVectors (:861-983); per pKF1 keypoint without MapPoint (and stereo when bOnlyStereo) one query
  std::vector<int> q1, qb, qe;
  {
    int a = 0, b = 0;
    while (a < k1->n_nodes && b < k2->n_nodes) {
      const uint32_t ka = k1->node_ids[a], kb = k2->node_ids[b];
      if (ka == kb) {
        for (int i = k1->node_off[a]; i < k1->node_off[a + 1]; ++i) {
          const uint32_t r = k1->idx[i];
          if (r >= (uint32_t)k1->n) { set_error("KF1 index out of range"); return B200ORB_EINVAL; }
          if (k1->has_mp && k1->has_mp[r]) continue;
          if (only_stereo && !(k1->uright[r] >= 0)) continue;
          q1.push_back((int)r); qb.push_back(k2->node_off[b]); qe.push_back(k2->node_off[b + 1]);
        }
        ++a; ++b;
      } else if (ka < kb) {
        a = (int)(std::lower_bound(k1->node_ids, k1->node_ids + k1->n_nodes, kb) - k1->node_ids);
      } else {
        b = (int)(std::lower_bound(k2->node_ids, k2->node_ids + k2->n_nodes, ka) - k2->node_ids);
      }
    }
  }
  const size_t nq = q1.size(), nfi = (size_t)k2->node_off[k2->n_nodes];
  if (nq == 0) return B200ORB_OK;
  for (size_t i = 0; i < nfi; ++i)
    if (k2->idx[i] >= (uint32_t)k2->n) { set_error("KF2 index out of range"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  const size_t n1 = k1->n, n2 = k2->n;
  const size_t need = nq * 12 + nfi * 4 + n1 * (32 + 16 + 4) + n2 * (32 + 16 + 4 + 1) + 64 * 256 + 8192;
  B200_CHECK(h->reserve(need));
  Carver cm(h->d_arena);
  int* d_q1 = cm.take<int>(nq); int* d_qb = cm.take<int>(nq); int* d_qe = cm.take<int>(nq); unsigned* d_fidx = cm.take<unsigned>(nfi);
  uint8_t* d_d1 = cm.take<uint8_t>(n1 * 32); uint8_t* d_d2 = cm.take<uint8_t>(n2 * 32);
  float *d_x1 = cm.take<float>(n1), *d_y1 = cm.take<float>(n1), *d_a1 = cm.take<float>(n1), *d_u1 = cm.take<float>(n1);
  float *d_x2 = cm.take<float>(n2), *d_y2 = cm.take<float>(n2), *d_a2 = cm.take<float>(n2), *d_u2 = cm.take<float>(n2);
  int* d_o2 = cm.take<int>(n2); uint8_t* d_b2 = cm.take<uint8_t>(n2);
  int* d_out = cm.take<int>(n1); int* d_nm = cm.take<int>(1);
  std::vector<uint8_t> blocked(n2, 0);
  if (k2->has_mp) memcpy(blocked.data(), k2->has_mp, n2);
  UP(d_q1, q1.data(), nq, int); UP(d_qb, qb.data(), nq, int); UP(d_qe, qe.data(), nq, int); UP(d_fidx, k2->idx, nfi, unsigned);
  UP(d_d1, k1->desc, n1 * 32, uint8_t); UP(d_d2, k2->desc, n2 * 32, uint8_t);
  UP(d_x1, k1->x, n1, float); UP(d_y1, k1->y, n1, float); UP(d_a1, k1->angle, n1, float); UP(d_u1, k1->uright, n1, float);
  UP(d_x2, k2->x, n2, float); UP(d_y2, k2->y, n2, float); UP(d_a2, k2->angle, n2, float); UP(d_u2, k2->uright, n2, float);
  UP(d_o2, k2->octave, n2, int); UP(d_b2, blocked.data(), n2, uint8_t);
  B200_CUDA(cudaMemsetAsync(d_out, 0xff, sizeof(int) * n1, h->stream));
  TriView t;
  memset(&t, 0, sizeof(t));
  t.q_idx1 = d_q1; t.f_beg = d_qb; t.f_end = d_qe; t.f_idx = d_fidx; t.desc1 = d_d1; t.desc2 = d_d2;
  t.x1 = d_x1; t.y1 = d_y1; t.ang1 = d_a1; t.ur1 = d_u1; t.x2 = d_x2; t.y2 = d_y2; t.ang2 = d_a2; t.ur2 = d_u2; t.oct2 = d_o2;
  t.blocked2 = d_b2;
  for (int i = 0; i < 9; ++i) t.F12[i] = F12[i];
  t.ex = ex; t.ey = ey;
  for (int i = 0; i < nlevels; ++i) { t.sf2[i] = scale_factors2[i]; t.sigma2_2[i] = level_sigma2_2[i]; }
  t.only_stereo = only_stereo; t.nq = (int)nq; t.n1 = (int)n1;
  B200_CUDA(cudaStreamSynchronize(h->stream));   // `blocked` (host vector) has been read by the upload
  k_tri_best<<<((int)nq + CAND_WARPS - 1) / CAND_WARPS, CAND_WARPS * 32, 0, h->stream>>>(t, d_out);
  k_tri_prune<<<1, 32, 0, h->stream>>>(t, check_ori, d_out, d_nm);
  h->launches += 2;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(matches12, d_out, sizeof(int) * n1, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(nmatches, d_nm, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

// smallest-to-largest thresholds t_k on ratio with  ceil(logf(ratio) / L) <= k  <=>  ratio <= t_k : MapPoint::PredictScale
// (src/MapPoint.cc:463-478) evaluates ceil(log(ratio)/mfLogScaleFactor) with the HOST's libm; bisecting the float line with
// that very function gives the device an exact, transcendental-free form of it
static void predict_scale_thresholds(float log_scale_factor, int nlevels, float* thr) {
  auto level = [&](float r) { return (int)std::ceil(std::log(r) / log_scale_factor); };
  for (int k = 0; k < nlevels - 1; ++k) {
    uint32_t lo = 0x00000001u, hi = 0x7f7fffffu;   // positive finite floats, ordered like their bit patterns
    float flo, fhi;
    memcpy(&fhi, &hi, 4);
    if (level(fhi) <= k) { thr[k] = fhi; continue; }
    memcpy(&flo, &lo, 4);
    if (level(flo) > k) { thr[k] = 0.f; continue; }
    while (hi - lo > 1) {   // invariant: level(lo) <= k < level(hi)
      const uint32_t mid = lo + (hi - lo) / 2;
      float fm;
      memcpy(&fm, &mid, 4);
      if (level(fm) <= k) lo = mid; else hi = mid;
    }
    memcpy(&thr[k], &lo, 4);
  }
}

int orbm_is_in_frustum(orbm_t* h, const OrbmFrame* f, const OrbmFrustumPoints* p, float viewing_cos_limit,
                       float log_scale_factor, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr,
                       int32_t* scale_level, float* view_cos) {
  if (!h || !f || !p || !in_view || !proj_x || !proj_y || !proj_xr || !scale_level || !view_cos) { set_error("null argument"); return B200ORB_EINVAL; }
  if (p->n < 0 || f->nlevels < 1 || f->nlevels > MAX_LEVELS || !(log_scale_factor > 0)) { set_error("bad argument"); return B200ORB_EINVAL; }
  if (p->n == 0) return B200ORB_OK;
  if (!p->xw || !p->normal || !p->min_dist || !p->max_dist) { set_error("null point array"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  const size_t n = p->n;
  B200_CHECK(h->reserve(n * (12 + 12 + 4 + 4 + 1 + 4 * 5) + 8192));
  Carver cm(h->d_arena);
  float* d_xw = cm.take<float>(n * 3); float* d_nrm = cm.take<float>(n * 3); float* d_mn = cm.take<float>(n); float* d_mx = cm.take<float>(n);
  uint8_t* d_in = cm.take<uint8_t>(n); float* d_px = cm.take<float>(n); float* d_py = cm.take<float>(n); float* d_pxr = cm.take<float>(n);
  int* d_lvl = cm.take<int>(n); float* d_vc = cm.take<float>(n);
  UP(d_xw, p->xw, n * 3, float); UP(d_nrm, p->normal, n * 3, float); UP(d_mn, p->min_dist, n, float); UP(d_mx, p->max_dist, n, float);
  B200_CUDA(cudaMemsetAsync(d_px, 0, 4 * n, h->stream)); B200_CUDA(cudaMemsetAsync(d_py, 0, 4 * n, h->stream));
  B200_CUDA(cudaMemsetAsync(d_pxr, 0, 4 * n, h->stream)); B200_CUDA(cudaMemsetAsync(d_lvl, 0, 4 * n, h->stream));
  B200_CUDA(cudaMemsetAsync(d_vc, 0, 4 * n, h->stream));
  FrustumView v;
  memset(&v, 0, sizeof(v));
  v.xw = d_xw; v.normal = d_nrm; v.min_dist = d_mn; v.max_dist = d_mx; v.n = (int)n;
  memcpy(v.Tcw, f->Tcw, 64);
  v.fx = f->fx; v.fy = f->fy; v.cx = f->cx; v.cy = f->cy; v.bf = f->bf;
  v.min_x = f->min_x; v.max_x = f->max_x; v.min_y = f->min_y; v.max_y = f->max_y;
  v.cos_limit = viewing_cos_limit; v.nlevels = f->nlevels;
  predict_scale_thresholds(log_scale_factor, f->nlevels, v.level_thr);
  k_is_in_frustum<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(v, d_in, d_px, d_py, d_pxr, d_lvl, d_vc);
  ++h->launches;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(in_view, d_in, n, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(proj_x, d_px, 4 * n, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(proj_y, d_py, 4 * n, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(proj_xr, d_pxr, 4 * n, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(scale_level, d_lvl, 4 * n, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(view_cos, d_vc, 4 * n, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int orbm_undistort_keypoints(orbm_t* h, const float* xy_in, int n, const float K[9], const float* dist, int ndist,
                             float* xy_out) {
  if (!h || !K || (n > 0 && (!xy_in || !xy_out)) || n < 0 || ndist < 0 || ndist > 5 || (ndist > 0 && !dist)) { set_error("bad argument"); return B200ORB_EINVAL; }
  if (n == 0) return B200ORB_OK;
  if (ndist == 0 || dist[0] == 0.0f) {   // src/Frame.cc:561-565: mvKeysUn = mvKeys
    memcpy(xy_out, xy_in, sizeof(float) * 2 * (size_t)n);
    return B200ORB_OK;
  }
  DeviceGuard g(h->device);
  B200_CHECK(h->reserve((size_t)n * 16 + 4096));
  Carver cm(h->d_arena);
  float* d_in = cm.take<float>((size_t)n * 2); float* d_out = cm.take<float>((size_t)n * 2);
  UP(d_in, xy_in, (size_t)n * 2, float);
  UndistortParams p;
  p.fx = K[0]; p.fy = K[4]; p.cx = K[2]; p.cy = K[5];
  for (int i = 0; i < 5; ++i) p.k[i] = i < ndist ? (double)dist[i] : 0.0;
  k_undistort_points<<<(n + 255) / 256, 256, 0, h->stream>>>(p, d_in, d_out, n);
  ++h->launches;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(xy_out, d_out, sizeof(float) * 2 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

static int bow_impl(orbm_t* h, const OrbmBow* kf, const OrbmBow* f, float nnratio, int check_ori, int kfkf, int32_t* f2kf,
                    int* nmatches) {
  if (!h || !kf || !f || !f2kf || !nmatches) { set_error("null argument"); return B200ORB_EINVAL; }
  if (kf->n < 0 || f->n < 0 || kf->n_nodes < 0 || f->n_nodes < 0 || f->n >= (1 << 20)) { set_error("bad counts"); return B200ORB_EINVAL; }
  *nmatches = 0;
  const int nout = kfkf ? kf->n : f->n;
  for (int j = 0; j < nout; ++j) f2kf[j] = -1;
  if (kf->n == 0 || f->n == 0 || kf->n_nodes == 0 || f->n_nodes == 0) return B200ORB_OK;
  // merge-join of the two FeatureVectors (:242-335): equal node ids pair up; queries in the reference's visiting order
  std::vector<int> q_kf, q_beg, q_end;
  {
    int a = 0, b = 0;
    while (a < kf->n_nodes && b < f->n_nodes) {
      const uint32_t ka = kf->node_ids[a], kb = f->node_ids[b];
      if (ka == kb) {
        for (int i = kf->node_off[a]; i < kf->node_off[a + 1]; ++i) {
          const uint32_t r = kf->idx[i];
          if (r >= (uint32_t)kf->n) { set_error("KF index out of range"); return B200ORB_EINVAL; }
          if (kf->valid && !kf->valid[r]) continue;
          q_kf.push_back((int)r); q_beg.push_back(f->node_off[b]); q_end.push_back(f->node_off[b + 1]);
        }
        ++a; ++b;
      } else if (ka < kb) {
        a = (int)(std::lower_bound(kf->node_ids, kf->node_ids + kf->n_nodes, kb) - kf->node_ids);
      } else {
        b = (int)(std::lower_bound(f->node_ids, f->node_ids + f->n_nodes, ka) - f->node_ids);
      }
    }
  }
  const size_t nq = q_kf.size(), nfi = (size_t)f->node_off[f->n_nodes];
  if (nq == 0) return B200ORB_OK;
  for (size_t i = 0; i < nfi; ++i)
    if (f->idx[i] >= (uint32_t)f->n) { set_error("F index out of range"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  const size_t need = nq * (12 + 4 * LCAP + 8) + nfi * 4 + (size_t)kf->n * (36 + 4) + (size_t)f->n * (36 + 4 + 1) + 64 * 256 + 4096;
  B200_CHECK(h->reserve(need));
  Carver cm(h->d_arena);
  int* d_qkf = cm.take<int>(nq); int* d_qb = cm.take<int>(nq); int* d_qe = cm.take<int>(nq);
  unsigned* d_fidx = cm.take<unsigned>(nfi);
  uint8_t* d_kd = cm.take<uint8_t>((size_t)kf->n * 32); uint8_t* d_fd = cm.take<uint8_t>((size_t)f->n * 32);
  float* d_ka = cm.take<float>(kf->n); float* d_fa = cm.take<float>(f->n);
  unsigned* d_list = cm.take<unsigned>(nq * LCAP); int* d_count = cm.take<int>(nq); int* d_acc = cm.take<int>(nq);
  int* d_out = cm.take<int>(nout); int* d_nm = cm.take<int>(1);
  uint8_t* d_fvalid = (kfkf && f->valid) ? cm.take<uint8_t>(f->n) : nullptr;
  if (d_fvalid) UP(d_fvalid, f->valid, f->n, uint8_t);
  UP(d_qkf, q_kf.data(), nq, int); UP(d_qb, q_beg.data(), nq, int); UP(d_qe, q_end.data(), nq, int);
  UP(d_fidx, f->idx, nfi, unsigned); UP(d_kd, kf->desc, (size_t)kf->n * 32, uint8_t); UP(d_fd, f->desc, (size_t)f->n * 32, uint8_t);
  UP(d_ka, kf->angle, kf->n, float); UP(d_fa, f->angle, f->n, float);
  BowQueries bq{d_qkf, d_qb, d_qe, d_fidx, d_kd, d_fd, d_ka, d_fa, d_fvalid, (int)nq, f->n, kf->n};
  ListView lsv{d_list, d_count};
  k_cand_bow<<<((int)nq + CAND_WARPS - 1) / CAND_WARPS, CAND_WARPS * 32, 0, h->stream>>>(bq, lsv);
  size_t smem;
  const int cmax = align_up(std::max(f->n, kf->n), 16);
  B200_CHECK(resolve_smem(cmax, &smem, (const void*)k_resolve_bow));
  k_resolve_bow<<<1, 32, smem, h->stream>>>(bq, nnratio, check_ori, kfkf, lsv, d_acc, d_out, d_nm, cmax);
  h->launches += 2;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(f2kf, d_out, sizeof(int) * (size_t)nout, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(nmatches, d_nm, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int orbm_search_by_bow(orbm_t* h, const OrbmBow* kf, const OrbmBow* f, float nnratio, int check_ori, int32_t* f2kf,
                       int* nmatches) {
  return bow_impl(h, kf, f, nnratio, check_ori, 0, f2kf, nmatches);
}

int orbm_search_by_bow_kf(orbm_t* h, const OrbmBow* kf1, const OrbmBow* kf2, float nnratio, int check_ori,
                          int32_t* matches12, int* nmatches) {
  return bow_impl(h, kf1, kf2, nnratio, check_ori, 1, matches12, nmatches);
}

}  // extern "C"
