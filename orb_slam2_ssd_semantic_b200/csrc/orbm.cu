// orbm.cu -- host side + C-ABI of the B200 ORB matcher (reference: src/ORBmatcher.cc, include/ORBmatcher.h).
#include <new>
#include <vector>

#include "orbm_host.h"

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
    off = align_up_sz(off + n * sizeof(T), 256);
    return p;
  }
};
}  // namespace

int launch_match_last(const MatchBatch& mb, const MatchCam& cam, int npairs, int cmax, cudaStream_t stream) {
  const size_t smem = (size_t)cmax * 5 + 16;
  if (smem > 160 * 1024) { set_error("too many keypoints per frame for the matcher's shared memory"); return B200ORB_EINVAL; }
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    B200_CUDA(cudaFuncSetAttribute(k_match_last, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k_match_last<<<npairs, MATCH_THREADS, smem, stream>>>(mb, cam, cmax);
  B200_CUDA(cudaGetLastError());
  return B200ORB_OK;
}

extern "C" {

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

int orbm_search_by_projection_last(orbm_t* h, const OrbmFrame* cur, const OrbmLast* last, float th, int mono,
                                   float nnratio, int check_ori, int32_t* cur2last, int* nmatches) {
  if (!h || !cur || !last || !cur2last || !nmatches) { set_error("null argument"); return B200ORB_EINVAL; }
  if (cur->n < 0 || last->n < 0 || cur->n >= (1 << 20)) { set_error("bad keypoint count"); return B200ORB_EINVAL; }
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
  const size_t need = (nc * (4 * 4 + 4 + 32 + 4 + 4 + 4) + nl * (12 + 1 + 4 + 4 + 32 + 4 + 8 * MATCH_K + 4 + 4) + 4096) + 64 * 256;
  B200_CHECK(h->reserve(need));
  Carver cv(h->d_arena);
  MatchBatch mb{};
  float* d_cx = cv.take<float>(nc); float* d_cy = cv.take<float>(nc); float* d_cang = cv.take<float>(nc);
  float* d_cur = cv.take<float>(nc); int* d_coct = cv.take<int>(nc); uint8_t* d_cdesc = cv.take<uint8_t>(nc * 32);
  int* d_cobs = cur->mp_obs ? cv.take<int>(nc) : nullptr;
  int* d_cn = cv.take<int>(1); float* d_cT = cv.take<float>(16);
  float* d_lxw = cv.take<float>(nl * 3); uint8_t* d_lvalid = cv.take<uint8_t>(nl); int* d_loct = cv.take<int>(nl);
  float* d_lang = cv.take<float>(nl); uint8_t* d_ldesc = cv.take<uint8_t>(nl * 32);
  int* d_lobs = last->mp_obs ? cv.take<int>(nl) : nullptr;
  int* d_ln = cv.take<int>(1); float* d_lT = cv.take<float>(16);
  int* d_out = cv.take<int>(nc); int* d_nm = cv.take<int>(1);
  unsigned long long* d_topk = cv.take<unsigned long long>(nl * MATCH_K);
  int* d_ncand = cv.take<int>(nl); int* d_gidx = cv.take<int>(nc); int* d_acc = cv.take<int>(nl);
  UP(d_cx, cur->x, nc, float); UP(d_cy, cur->y, nc, float); UP(d_cang, cur->angle, nc, float);
  UP(d_cur, cur->uright, nc, float); UP(d_coct, cur->octave, nc, int); UP(d_cdesc, cur->desc, nc * 32, uint8_t);
  if (d_cobs) UP(d_cobs, cur->mp_obs, nc, int);
  const int cn = cur->n, ln = last->n;
  UP(d_cn, &cn, 1, int); UP(d_cT, cur->Tcw, 16, float);
  UP(d_lxw, last->xw, nl * 3, float); UP(d_lvalid, last->valid, nl, uint8_t); UP(d_loct, last->octave, nl, int);
  UP(d_lang, last->angle, nl, float); UP(d_ldesc, last->mp_desc, nl * 32, uint8_t);
  if (d_lobs) UP(d_lobs, last->mp_obs, nl, int);
  UP(d_ln, &ln, 1, int); UP(d_lT, last->Tcw, 16, float);
  mb.cx = d_cx; mb.cy = d_cy; mb.cang = d_cang; mb.curight = d_cur; mb.coct = d_coct; mb.cdesc = d_cdesc;
  mb.cobs = d_cobs; mb.cn = d_cn; mb.cTcw = d_cT; mb.cstride = nc;
  mb.lxw = d_lxw; mb.lvalid = d_lvalid; mb.loct = d_loct; mb.lang = d_lang; mb.ldesc = d_ldesc; mb.lobs = d_lobs;
  mb.ln = d_ln; mb.lTcw = d_lT; mb.lstride = nl;
  mb.cur2last = d_out; mb.nmatch = d_nm; mb.topk = d_topk; mb.ncand = d_ncand; mb.grididx = d_gidx; mb.accepted = d_acc;
  B200_CHECK(launch_match_last(mb, cam, 1, align_up((int)nc, 16), h->stream));
  ++h->launches;
  B200_CUDA(cudaMemcpyAsync(cur2last, d_out, sizeof(int) * nc, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(nmatches, d_nm, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

}  // extern "C"
