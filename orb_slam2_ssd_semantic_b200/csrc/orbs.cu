// orbs.cu -- batched stream pipeline (north_star's "batched many-frame mode"): per frame t of a batch the device
// does what Tracking::GrabImageRGBD -> Frame::Frame(RGB-D) -> TrackWithMotionModel's SearchByProjection do on the
// CPU (src/Tracking.cc:331-375,1324-1352; src/Frame.cc:176-240), with no host round trip between the stages.
#include <new>

#include "orbm_host.h"
#include "orbx_host.h"

using namespace b200;

namespace b200 {

struct GlueOut {     // SoA per frame, stride cap
  float *x, *y, *ang, *uright, *depth, *xw;
  int* oct;
  uint8_t* valid;
};

// Frame::ComputeStereoFromRGBD (src/Frame.cc:850-871) + Frame::UnprojectStereo (:879-899) for every keypoint,
// and the AoS -> SoA split the matcher kernel reads.
__global__ void __launch_bounds__(256) k_frame_glue(const OrbxKeyPoint* __restrict__ kps, const int* __restrict__ nkp,
                                                    int cap, const float* __restrict__ depth,
                                                    const uint16_t* __restrict__ depth16,
                                                    const uint16_t* __restrict__ kpd16, float depth_factor, int rows,
                                                    int cols, const float* __restrict__ Tcw, float fx, float fy,
                                                    float cx, float cy, float bf, GlueOut o) {
  const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nkp[f]) return;
  const size_t g = (size_t)f * cap + i;
  const OrbxKeyPoint kp = kps[g];
  const float* T = Tcw + (size_t)f * 16;
  const float u = kp.x, v = kp.y;
  // imDepth.at<float>(v,u) (src/Frame.cc:863).  With depth16 the CV_16U sensor image is read in place -- possibly
  // straight out of pinned host memory over PCIe, one 2-byte read per keypoint instead of uploading 0.6 MB per frame
  // -- and converted like convertTo(CV_32F, mDepthMapFactor) does (src/Tracking.cc:366-367).
  const size_t di = (size_t)f * rows * cols + (size_t)(int)v * cols + (int)u;
  // kpd16: the same CV_16U pixel, already fetched per keypoint by k_depth_prefetch
  const float d = kpd16 ? __fmul_rn((float)kpd16[g], depth_factor)
                        : (depth16 ? __fmul_rn((float)depth16[di], depth_factor) : depth[di]);
  o.x[g] = u; o.y[g] = v; o.ang[g] = kp.angle; o.oct[g] = kp.octave;
  float ur = -1.f, dd = -1.f;
  uint8_t ok = 0;
  if (d > 0) {
    dd = d;
    ur = __fsub_rn(u, __fdiv_rn(bf, d));
    const float invfx = __fdiv_rn(1.0f, fx), invfy = __fdiv_rn(1.0f, fy);   // src/Frame.cc:215-216
    const float x = __fmul_rn(__fmul_rn(__fsub_rn(u, cx), d), invfx);
    const float y = __fmul_rn(__fmul_rn(__fsub_rn(v, cy), d), invfy);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float Rwc[3] = {T[0 * 4 + r], T[1 * 4 + r], T[2 * 4 + r]};   // mRwc = mRcw.t()
      // mOw = -mRwc*mtcw (src/Frame.cc:373 -- this fork multiplies the EVALUATED transpose, so the product takes the
      // small-matrix gemm path: float accumulation left to right, alpha = -1)
      const float ow = -__fadd_rn(__fadd_rn(__fmul_rn(Rwc[0], T[3]), __fmul_rn(Rwc[1], T[7])), __fmul_rn(Rwc[2], T[11]));
      o.xw[g * 3 + r] = gemm3(Rwc, x, y, d, ow);
    }
    ok = 1;
  }
  o.uright[g] = ur; o.depth[g] = dd; o.valid[g] = ok;
}

// Depth under the selected keypoints, fetched as soon as the selection exists (the final KeyPoint coordinates are the
// level coordinates times the level's scale factor, src/ORBextractor.cc:1104-1110, exactly as k_orient_desc forms them)
// -- on a side stream, so the PCIe round trips of the in-place host reads overlap blur and descriptors.
// The reads are latency-bound on the host link (~2 us each, a bounded number in flight), not on the SMs, and every
// read that is queued behind that bound sits in the SM -> L2 request path in front of the HBM traffic of the kernels
// running beside it.  Measured on a B200, 256 frames x 2000 keypoints, blur running beside the prefetch:
//   one thread per item (590 k reads queued): blur 0.32 -> 1.43 ms;   64 CTAs x 256 thr x 4 reads: 1.45 ms;
//   16 CTAs: 0.60 ms;   4 CTAs (4096 reads in flight): 0.34 ms, prefetch 1.06 ms;   1 CTA: prefetch 2.9 ms.
// Whole e2e step (bench.py): 46.1 k frames/s before, 53.9 k with 4 CTAs, 54.8 k with 8.
// So: DEPTH_PF_CTAS small CTAs walk the (frame, keypoint) items with DEPTH_PF_UNROLL reads in flight per thread.
constexpr int DEPTH_PF_THREADS = 256, DEPTH_PF_UNROLL = 4, DEPTH_PF_CTAS = 8;
__global__ void __launch_bounds__(DEPTH_PF_THREADS) k_depth_prefetch(LevelTab lt, const unsigned* __restrict__ sel,
                                                        const int* __restrict__ selcnt, int sel_per_frame, int cap,
                                                        const uint16_t* __restrict__ depth16, int rows, int cols, int f0,
                                                        int nframes, uint16_t* __restrict__ kpd16) {
  const long long total = (long long)nframes * cap, stride = (long long)gridDim.x * DEPTH_PF_THREADS;
  for (long long base = (long long)blockIdx.x * DEPTH_PF_THREADS + threadIdx.x; base < total; base += stride * DEPTH_PF_UNROLL) {
    const uint16_t* src[DEPTH_PF_UNROLL];
    long long dst[DEPTH_PF_UNROLL];
#pragma unroll
    for (int k = 0; k < DEPTH_PF_UNROLL; ++k) {
      src[k] = nullptr;
      const long long it = base + k * stride;
      if (it >= total) continue;
      const int f = f0 + (int)(it / cap), j = (int)(it % cap);
      const int* sc = selcnt + (size_t)f * lt.nlevels;
      int l = -1, idx = 0, acc = 0;
      for (int q = 0; q < lt.nlevels; ++q) {
        const int c = sc[q];
        if (l < 0 && j < acc + c) { l = q; idx = j - acc; }
        acc += c;
      }
      if (l < 0) continue;
      const unsigned pk = sel[(size_t)f * sel_per_frame + lt.sel_off[l] + idx];
      const float s = lt.sf[l];
      const float u = (l != 0) ? __fmul_rn((float)kp_x(pk), s) : (float)kp_x(pk);
      const float v = (l != 0) ? __fmul_rn((float)kp_y(pk), s) : (float)kp_y(pk);
      src[k] = depth16 + (size_t)f * rows * cols + (size_t)(int)v * cols + (int)u;
      dst[k] = (long long)f * cap + j;
    }
    uint16_t val[DEPTH_PF_UNROLL];
#pragma unroll
    for (int k = 0; k < DEPTH_PF_UNROLL; ++k) val[k] = src[k] ? *src[k] : (uint16_t)0;
#pragma unroll
    for (int k = 0; k < DEPTH_PF_UNROLL; ++k)
      if (src[k]) kpd16[dst[k]] = val[k];
  }
}

// imDepth.convertTo(CV_32F, mDepthMapFactor) (src/Tracking.cc:366-367): float(u16) * factor, 4 pixels per thread
__global__ void k_depth_u16_to_f32(const ushort4* __restrict__ src, float4* __restrict__ dst, float factor, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const ushort4 v = src[i];
  dst[i] = make_float4(__fmul_rn((float)v.x, factor), __fmul_rn((float)v.y, factor), __fmul_rn((float)v.z, factor),
                       __fmul_rn((float)v.w, factor));
}

__global__ void k_fill_i32(int* p, int v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace b200

struct orbs {
  OrbsParams prm{};
  int device = 0;
  orbx* ex = nullptr;
  long long launches = 0;
  int rows = 0, cols = 0, maxF = 0, cap = 0, lastF = 0;
  float *d_x = nullptr, *d_y = nullptr, *d_ang = nullptr, *d_ur = nullptr, *d_dep = nullptr, *d_xw = nullptr;
  int* d_oct = nullptr;
  uint8_t* d_valid = nullptr;
  int *d_c2l = nullptr, *d_nm = nullptr, *d_count = nullptr, *d_gidx = nullptr, *d_goff = nullptr, *d_acc = nullptr;
  unsigned* d_list = nullptr;
  uint8_t* d_gray = nullptr;
  float *d_depth = nullptr, *d_T = nullptr;
  uint16_t* d_depth16 = nullptr;
  bool full_depth_valid = false, force_full_depth_upload = false;
  int chunk_frames = 128;   // frames per upload chunk of the host-buffer entries
  cudaStream_t copy_stream = nullptr;
  cudaStream_t gather_stream = nullptr;   // in-place depth reads under the keypoints (sparse mode)
  cudaEvent_t sel_ev = nullptr, gather_ev = nullptr;
  cudaEvent_t compute_done_ev = nullptr;   // recorded after the last kernel of a host-buffer batch (before its downloads)
  orbs* chain_prev = nullptr;              // orbs_chain_after: the next batch's kernels wait for that handle's kernels
  uint16_t* d_kpd16 = nullptr;            // [maxF][cap] depth under the keypoints
  const uint16_t* sparse_d16 = nullptr;   // device alias of the caller's page-locked depth during a sparse call
  cudaEvent_t chunk_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool have = false;
  void free_bufs() {
    auto F = [](void* p) { if (p) cudaFree(p); };
    F(d_x); F(d_y); F(d_ang); F(d_ur); F(d_dep); F(d_xw); F(d_oct); F(d_valid); F(d_c2l); F(d_nm); F(d_count);
    F(d_gidx); F(d_goff); F(d_acc); F(d_list); F(d_gray); F(d_depth); F(d_T); F(d_depth16); F(d_kpd16);
    d_depth16 = nullptr; d_kpd16 = nullptr;
    d_x = d_y = d_ang = d_ur = d_dep = d_xw = nullptr; d_oct = nullptr; d_valid = nullptr;
    d_c2l = d_nm = d_count = d_gidx = d_goff = d_acc = nullptr; d_list = nullptr; d_gray = nullptr; d_depth = d_T = nullptr;
  }
  ~orbs() {
    DeviceGuard g(device);
    free_bufs();
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (gather_stream) cudaStreamDestroy(gather_stream);
    if (sel_ev) cudaEventDestroy(sel_ev);
    if (compute_done_ev) cudaEventDestroy(compute_done_ev);
    if (gather_ev) cudaEventDestroy(gather_ev);
    for (cudaEvent_t e : chunk_ev) if (e) cudaEventDestroy(e);
    delete ex;
  }
  int ensure(int r, int c, int F, bool host_inputs) {
    B200_CHECK(ex->ensure_geometry(r, c, F));
    if (!(r == rows && c == cols && F <= maxF && ex->cap == cap)) {
      B200_CUDA(cudaStreamSynchronize(ex->stream));
      free_bufs();
      rows = r; cols = c; maxF = std::max(F, ex->maxF); cap = ex->cap;
      const size_t n = (size_t)maxF * cap;
      B200_CUDA(cudaMalloc(&d_x, n * 4)); B200_CUDA(cudaMalloc(&d_y, n * 4)); B200_CUDA(cudaMalloc(&d_ang, n * 4));
      B200_CUDA(cudaMalloc(&d_ur, n * 4)); B200_CUDA(cudaMalloc(&d_dep, n * 4)); B200_CUDA(cudaMalloc(&d_xw, n * 12));
      B200_CUDA(cudaMalloc(&d_oct, n * 4)); B200_CUDA(cudaMalloc(&d_valid, n));
      B200_CUDA(cudaMalloc(&d_c2l, n * 4)); B200_CUDA(cudaMalloc(&d_nm, (size_t)maxF * 4));
      B200_CUDA(cudaMalloc(&d_count, n * 4)); B200_CUDA(cudaMalloc(&d_gidx, n * 4)); B200_CUDA(cudaMalloc(&d_acc, n * 4));
      B200_CUDA(cudaMalloc(&d_goff, (size_t)maxF * (GRID_CELLS + 1) * 4));
      B200_CUDA(cudaMalloc(&d_list, n * 4 * LCAP));
    }
    if (host_inputs && !d_gray) {
      B200_CUDA(cudaMalloc(&d_gray, (size_t)maxF * rows * cols));
      B200_CUDA(cudaMalloc(&d_depth, (size_t)maxF * rows * cols * 4));
      B200_CUDA(cudaMalloc(&d_T, (size_t)maxF * 64));
      B200_CUDA(cudaMalloc(&d_depth16, (size_t)maxF * rows * cols * 2));
      B200_CUDA(cudaMalloc(&d_kpd16, (size_t)maxF * cap * 2));
    }
    return B200ORB_OK;
  }
  int run(const uint8_t* dg, const float* dd, const float* dT, int F, const uint16_t* dd16 = nullptr, float factor = 0.f,
          bool extracted = false, const uint16_t* kpd16 = nullptr) {
    cudaStream_t st = ex->stream;
    if (!extracted) B200_CHECK(ex->run(dg, cols, (size_t)rows * cols, F));
    GlueOut go{d_x, d_y, d_ang, d_ur, d_dep, d_xw, d_oct, d_valid};
    k_frame_glue<<<dim3((cap + 255) / 256, F), 256, 0, st>>>(ex->d_kps, ex->d_n, cap, dd, dd16, kpd16, factor, rows, cols, dT, prm.fx, prm.fy,
                                                            prm.cx, prm.cy, prm.bf, go);
    ++launches;
    B200_CHECK(ex->prof_mark(ST_GLUE + 1));
    // frame 0 of the batch has no predecessor: all -1, 0 matches
    k_fill_i32<<<(cap + 255) / 256, 256, 0, st>>>(d_c2l, -1, (size_t)cap);
    ++launches;
    B200_CUDA(cudaMemsetAsync(d_nm, 0, 4, st));
    if (F > 1) {
      MatchCam cam;
      memset(&cam, 0, sizeof(cam));
      cam.fx = prm.fx; cam.fy = prm.fy; cam.cx = prm.cx; cam.cy = prm.cy; cam.bf = prm.bf;
      cam.b = prm.bf / prm.fx;                       // mb = mbf/fx, src/Frame.cc:218
      cam.min_x = 0.f; cam.max_x = (float)cols;      // no distortion: ComputeImageBounds, src/Frame.cc:617-623
      cam.min_y = 0.f; cam.max_y = (float)rows;
      for (int l = 0; l < prm.orb.nlevels; ++l) cam.sf[l] = ex->sf[l];
      cam.th = prm.th; cam.nnratio = prm.nnratio; cam.mono = 0; cam.check_ori = prm.check_ori;
      cam.nlevels = prm.orb.nlevels;
      cam.last_obs_default = 1;                      // last-frame points behave like mapped points (Observations()>0)
      const size_t c = cap;
      CurView cv;
      memset(&cv, 0, sizeof(cv));
      cv.x = d_x + c; cv.y = d_y + c; cv.ang = d_ang + c; cv.uright = d_ur + c; cv.oct = d_oct + c;
      cv.desc = ex->d_desc + c * 32; cv.obs = nullptr; cv.n = ex->d_n + 1; cv.Tcw = dT + 16; cv.stride = c;
      LastView lv{d_xw, d_valid, d_oct, d_ang, ex->d_desc, nullptr, ex->d_n, dT, c};
      B200_CHECK(launch_match_last(cv, lv, cam, F - 1, d_goff, d_gidx, d_list, d_count, d_acc, d_c2l + c, d_nm + 1,
                                   align_up(cap, 16), cap, st, &launches));
    }
    B200_CHECK(ex->prof_mark(ST_MATCH + 1));
    lastF = F;
    have = true;
    return B200ORB_OK;
  }
};

extern "C" {

int orbs_create(const OrbsParams* p, int device, orbs_t** out) {
  if (!p || !out) { set_error("null argument"); return B200ORB_EINVAL; }
  *out = nullptr;
  orbx_t* ex = nullptr;
  B200_CHECK(orbx_create(&p->orb, device, &ex));
  orbs* h = new (std::nothrow) orbs();
  if (!h) { orbx_destroy(ex); set_error("out of host memory"); return B200ORB_EINVAL; }
  h->prm = *p;
  h->device = device;
  h->ex = ex;
  *out = h;
  return B200ORB_OK;
}
void orbs_destroy(orbs_t* h) { delete h; }

int orbs_track_batch_device(orbs_t* h, const uint8_t* d_gray, const float* d_depth, const float* d_Tcw, int nframes,
                            int rows, int cols) {
  if (!h || !d_gray || !d_depth || !d_Tcw || nframes <= 0 || rows <= 0 || cols <= 0) { set_error("bad argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  B200_CHECK(h->ensure(rows, cols, nframes, false));
  return h->run(d_gray, d_depth, d_Tcw, nframes);
}

int orbs_device_results(orbs_t* h, const OrbxKeyPoint** d_kps, const uint8_t** d_desc, const int32_t** d_nkp,
                        const int32_t** d_cur2last, const int32_t** d_nmatch, int* cap) {
  if (!h || !h->have) { set_error("no results"); return B200ORB_EINVAL; }
  if (d_kps) *d_kps = h->ex->d_kps;
  if (d_desc) *d_desc = h->ex->d_desc;
  if (d_nkp) *d_nkp = h->ex->d_n;
  if (d_cur2last) *d_cur2last = h->d_c2l;
  if (d_nmatch) *d_nmatch = h->d_nm;
  if (cap) *cap = h->cap;
  return B200ORB_OK;
}

// Frame-glue outputs of frame `frame` of the last batch (Frame::ComputeStereoFromRGBD + UnprojectStereo per keypoint):
// mvuRight, mvDepth, the world point of the keypoint under the frame's pose, and whether it has depth.
int orbs_read_frame_glue(orbs_t* h, int frame, float* uright, float* depth, float* xw, uint8_t* valid, int cap) {
  if (!h || !h->have || frame < 0 || cap < h->cap) { set_error("bad argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  const size_t o = (size_t)frame * h->cap, n = (size_t)h->cap;
  cudaStream_t st = h->ex->stream;
  if (uright) B200_CUDA(cudaMemcpyAsync(uright, h->d_ur + o, 4 * n, cudaMemcpyDeviceToHost, st));
  if (depth) B200_CUDA(cudaMemcpyAsync(depth, h->d_dep + o, 4 * n, cudaMemcpyDeviceToHost, st));
  if (xw) B200_CUDA(cudaMemcpyAsync(xw, h->d_xw + o * 3, 12 * n, cudaMemcpyDeviceToHost, st));
  if (valid) B200_CUDA(cudaMemcpyAsync(valid, h->d_valid + o, n, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  return B200ORB_OK;
}

// orbx::after_select hook of the sparse-depth mode: the keypoints of frames [f0, f0 + F) are selected -> fetch the depth
// under them on the side stream (the stream of the extractor goes on with blur and descriptors meanwhile)
static int orbs_after_select(void* ctx, int f0, int F) {
  orbs* h = static_cast<orbs*>(ctx);
  orbx* ex = h->ex;
  B200_CUDA(cudaEventRecord(h->sel_ev, ex->stream));
  B200_CUDA(cudaStreamWaitEvent(h->gather_stream, h->sel_ev, 0));
  const long long items = (long long)F * h->cap;
  int ctas = DEPTH_PF_CTAS;
  if (const char* e = getenv("B200ORB_PF_CTAS")) ctas = std::max(1, atoi(e));   // tuning knob (see k_depth_prefetch)
  const int pf_grid = (int)std::min<long long>(ctas, (items + DEPTH_PF_THREADS - 1) / DEPTH_PF_THREADS);
  k_depth_prefetch<<<pf_grid, DEPTH_PF_THREADS, 0, h->gather_stream>>>(ex->ltab, ex->d_sel, ex->d_selcnt, ex->sel_per_frame, h->cap,
                                                                       h->sparse_d16, h->rows, h->cols, f0, F, h->d_kpd16);
  ++h->launches;
  B200_CUDA(cudaEventRecord(h->gather_ev, h->gather_stream));
  return B200ORB_OK;
}

static int track_batch_host(orbs_t* h, const uint8_t* gray, const float* depth, const uint16_t* depth16, float factor,
                            const float* Tcw, int nframes, int rows, int cols, OrbxKeyPoint* kps, uint8_t* desc,
                            int32_t* nkp, int32_t* cur2last, int32_t* nmatch, int cap, bool wait = true) {
  if (!h || !gray || (!depth && !depth16) || !Tcw || !kps || !desc || !nkp || !cur2last || !nmatch || nframes <= 0 ||
      rows <= 0 || cols <= 0) {
    set_error("bad argument");
    return B200ORB_EINVAL;
  }
  DeviceGuard g(h->device);
  B200_CHECK(h->ensure(rows, cols, nframes, true));
  if (cap < h->cap) { set_error("cap %d < required %d (orbx_max_keypoints)", cap, h->cap); return B200ORB_ECAP; }
  cudaStream_t st = h->ex->stream;
  const size_t px = (size_t)rows * cols, F = nframes;
  // gray images go up in chunks on a copy stream while the extractor already works on the chunks that have arrived
  // (PCIe and SMs overlap); everything after extraction runs on the whole batch.
  if (!h->copy_stream) {
    B200_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : h->chunk_ev) B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const int nchunk = (int)std::min<size_t>(7, (F + h->chunk_frames - 1) / h->chunk_frames);   // >= 64-frame chunks keep each launch wide enough
  // Page-locked CV_16U depth (cudaHostAlloc / cudaHostRegister, e.g. torch.pin_memory): the only depth pixels the
  // tracking path ever reads are the ones under the keypoints (Frame::ComputeStereoFromRGBD), so the 0.6 MB/frame
  // images are never uploaded; the pixels are read in place over PCIe as soon as a chunk's keypoints are selected.
  const uint16_t* d16_sparse = nullptr;
  if (depth16 && !h->force_full_depth_upload) {
    cudaPointerAttributes at;
    void* dev_alias = nullptr;
    if (cudaPointerGetAttributes(&at, depth16) == cudaSuccess && at.type == cudaMemoryTypeHost &&
        cudaHostGetDevicePointer(&dev_alias, const_cast<uint16_t*>(depth16), 0) == cudaSuccess && dev_alias)
      d16_sparse = reinterpret_cast<const uint16_t*>(dev_alias);
    else
      cudaGetLastError();
  }
  if (d16_sparse && !h->gather_stream) {
    B200_CUDA(cudaStreamCreateWithFlags(&h->gather_stream, cudaStreamNonBlocking));
    B200_CUDA(cudaEventCreateWithFlags(&h->sel_ev, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&h->gather_ev, cudaEventDisableTiming));
  }
  h->sparse_d16 = d16_sparse;
  h->ex->after_select = d16_sparse ? &orbs_after_select : nullptr;
  h->ex->after_select_ctx = h;
  {
    // the copy stream may only overwrite d_gray once the previous call's kernels are done with it
    B200_CUDA(cudaEventRecord(h->chunk_ev[7], st));
    B200_CUDA(cudaStreamWaitEvent(h->copy_stream, h->chunk_ev[7], 0));
    // orbs_chain_after: this batch's kernels start when the other handle's kernels are done -- the uploads enqueued
    // below (copy stream) and that handle's downloads run meanwhile, and kernels of two batches never share the SMs
    if (h->chain_prev && h->chain_prev->compute_done_ev) B200_CUDA(cudaStreamWaitEvent(st, h->chain_prev->compute_done_ev, 0));
    h->chain_prev = nullptr;
    for (int c = 0; c < nchunk; ++c) {
      const size_t f0 = F * c / nchunk, f1 = F * (c + 1) / nchunk;
      B200_CUDA(cudaMemcpyAsync(h->d_gray + px * f0, gray + px * f0, px * (f1 - f0), cudaMemcpyHostToDevice, h->copy_stream));
      B200_CUDA(cudaEventRecord(h->chunk_ev[c], h->copy_stream));
      B200_CUDA(cudaStreamWaitEvent(st, h->chunk_ev[c], 0));
      B200_CHECK(h->ex->run(h->d_gray + px * f0, cols, px, (int)(f1 - f0), (int)f0));
    }
  }
  h->ex->after_select = nullptr;
  h->full_depth_valid = true;
  if (d16_sparse) {
    h->full_depth_valid = false;
    B200_CUDA(cudaStreamWaitEvent(st, h->gather_ev, 0));   // the last chunk's prefetch (the side stream is in order)
  } else if (depth16) {
    if ((px * F) % 4) { set_error("u16 depth path needs rows*cols*nframes to be a multiple of 4"); return B200ORB_EINVAL; }
    B200_CUDA(cudaMemcpyAsync(h->d_depth16, depth16, px * F * 2, cudaMemcpyHostToDevice, st));
    const size_t n4 = px * F / 4;
    k_depth_u16_to_f32<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const ushort4*>(h->d_depth16),
                                                                     reinterpret_cast<float4*>(h->d_depth), factor, n4);
    ++h->launches;
  } else {
    B200_CUDA(cudaMemcpyAsync(h->d_depth, depth, px * F * 4, cudaMemcpyHostToDevice, st));
  }
  B200_CUDA(cudaMemcpyAsync(h->d_T, Tcw, F * 64, cudaMemcpyHostToDevice, st));
  B200_CHECK(h->run(h->d_gray, h->d_depth, h->d_T, nframes, d16_sparse, factor, /*extracted=*/true, d16_sparse ? h->d_kpd16 : nullptr));
  if (!h->compute_done_ev) B200_CUDA(cudaEventCreateWithFlags(&h->compute_done_ev, cudaEventDisableTiming));
  B200_CUDA(cudaEventRecord(h->compute_done_ev, st));
  const size_t hc = h->cap;
  if ((size_t)cap == hc) {
    B200_CUDA(cudaMemcpyAsync(kps, h->ex->d_kps, sizeof(OrbxKeyPoint) * hc * F, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(desc, h->ex->d_desc, 32 * hc * F, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(cur2last, h->d_c2l, 4 * hc * F, cudaMemcpyDeviceToHost, st));
  } else {
    B200_CUDA(cudaMemcpy2DAsync(kps, sizeof(OrbxKeyPoint) * cap, h->ex->d_kps, sizeof(OrbxKeyPoint) * hc,
                                sizeof(OrbxKeyPoint) * hc, F, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpy2DAsync(desc, (size_t)32 * cap, h->ex->d_desc, 32 * hc, 32 * hc, F, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpy2DAsync(cur2last, (size_t)4 * cap, h->d_c2l, 4 * hc, 4 * hc, F, cudaMemcpyDeviceToHost, st));
  }
  B200_CUDA(cudaMemcpyAsync(nkp, h->ex->d_n, 4 * F, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaMemcpyAsync(nmatch, h->d_nm, 4 * F, cudaMemcpyDeviceToHost, st));
  if (wait) B200_CUDA(cudaStreamSynchronize(st));
  return B200ORB_OK;
}

int orbs_track_batch(orbs_t* h, const uint8_t* gray, const float* depth, const float* Tcw, int nframes, int rows,
                     int cols, OrbxKeyPoint* kps, uint8_t* desc, int32_t* nkp, int32_t* cur2last, int32_t* nmatch,
                     int cap) {
  return track_batch_host(h, gray, depth, nullptr, 0.f, Tcw, nframes, rows, cols, kps, desc, nkp, cur2last, nmatch, cap);
}

int orbs_track_batch_u16(orbs_t* h, const uint8_t* gray, const uint16_t* depth_u16, float depth_factor, const float* Tcw,
                         int nframes, int rows, int cols, OrbxKeyPoint* kps, uint8_t* desc, int32_t* nkp,
                         int32_t* cur2last, int32_t* nmatch, int cap) {
  return track_batch_host(h, gray, nullptr, depth_u16, depth_factor, Tcw, nframes, rows, cols, kps, desc, nkp, cur2last,
                          nmatch, cap);
}

int orbs_submit_batch_u16(orbs_t* h, const uint8_t* gray, const uint16_t* depth_u16, float depth_factor, const float* Tcw,
                          int nframes, int rows, int cols, OrbxKeyPoint* kps, uint8_t* desc, int32_t* nkp,
                          int32_t* cur2last, int32_t* nmatch, int cap) {
  return track_batch_host(h, gray, nullptr, depth_u16, depth_factor, Tcw, nframes, rows, cols, kps, desc, nkp, cur2last,
                          nmatch, cap, /*wait=*/false);
}

int orbs_device_inputs(orbs_t* h, const uint8_t** d_gray, const float** d_depth) {
  if (!h || !h->d_gray) { set_error("no host-buffer call yet"); return B200ORB_EINVAL; }
  if (d_gray) *d_gray = h->d_gray;
  if (d_depth) *d_depth = h->full_depth_valid ? h->d_depth : nullptr;   // NULL after a sparse (zero-copy) depth call
  return B200ORB_OK;
}

// 1: always upload the whole CV_16U depth batch (device copy available through orbs_device_inputs); 0 (default): read
// the depth under the keypoints in place when the host buffer is page-locked.
int orbs_set_full_depth_upload(orbs_t* h, int on) {
  if (!h) return B200ORB_EINVAL;
  h->force_full_depth_upload = on != 0;
  return B200ORB_OK;
}

// convertTo(CV_32F, factor) of n CV_16U pixels already in HBM (n multiple of 4), on `stream` (cudaStream_t)
int b200orb_depth_u16_to_f32_device(const uint16_t* d_src, float* d_dst, size_t n, float factor, void* stream) {
  if (!d_src || !d_dst || (n % 4)) { set_error("bad argument (n must be a multiple of 4)"); return B200ORB_EINVAL; }
  k_depth_u16_to_f32<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const ushort4*>(d_src), reinterpret_cast<float4*>(d_dst), factor, n / 4);
  B200_CUDA(cudaGetLastError());
  return B200ORB_OK;
}

// frames per upload chunk of the host-buffer entries (default 128; at most 7 chunks per call): smaller chunks start the
// extraction earlier, larger ones keep the launches wider
int orbs_set_chunk_frames(orbs_t* h, int frames) {
  if (!h || frames <= 0) { set_error("bad argument"); return B200ORB_EINVAL; }
  h->chunk_frames = frames;
  return B200ORB_OK;
}

// The next host-buffer batch submitted to `h` starts its kernels only when the kernels of `prev`'s last submitted batch
// are done (its uploads are not held back).  Two handles submitted alternately with this form a 3-stage pipeline:
// upload of batch k+1 | kernels of batch k | download of batch k-1.
int orbs_chain_after(orbs_t* h, orbs_t* prev) {
  if (!h || !prev || h == prev || h->device != prev->device) { set_error("bad argument"); return B200ORB_EINVAL; }
  h->chain_prev = prev;
  return B200ORB_OK;
}

int orbs_sync(orbs_t* h) {
  if (!h) return B200ORB_EINVAL;
  DeviceGuard g(h->device);
  B200_CUDA(cudaStreamSynchronize(h->ex->stream));
  return B200ORB_OK;
}
void* orbs_stream(orbs_t* h) { return h ? (void*)h->ex->stream : nullptr; }
long long orbs_launch_count(const orbs_t* h) { return h ? h->launches + h->ex->launches : 0; }
orbx_t* orbs_extractor(orbs_t* h) { return h ? h->ex : nullptr; }

}  // extern "C"
