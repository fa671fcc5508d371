// orbm_match.cuh -- the SearchByProjection(CurrentFrame, LastFrame) kernel (src/ORBmatcher.cc:1578-1724).
#pragma once
#include "orbm_kernels.cuh"

namespace b200 {

struct LastQuery {   // per-query projection state shared by phase B and the re-walk of phase C
  QueryGeom g;
  uint4 d0, d1;
};

// Projection + window set-up of one last-frame MapPoint (:1616-1644).
__device__ __forceinline__ bool setup_last_query(const MatchCam& cam, const float* Rt /*cur Tcw row-major 16*/,
                                                 bool bForward, bool bBackward, const float* __restrict__ xw,
                                                 int octave, const uint8_t* __restrict__ desc, LastQuery& q) {
  const float R0[3] = {Rt[0], Rt[1], Rt[2]}, R1[3] = {Rt[4], Rt[5], Rt[6]}, R2[3] = {Rt[8], Rt[9], Rt[10]};
  const float xc = gemm3(R0, xw[0], xw[1], xw[2], Rt[3]);
  const float yc = gemm3(R1, xw[0], xw[1], xw[2], Rt[7]);
  const float zc = gemm3(R2, xw[0], xw[1], xw[2], Rt[11]);
  const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)zc));   // const float invzc = 1.0/x3Dc.at<float>(2)
  if (invzc < 0) return false;
  const float u = __fadd_rn(__fmul_rn(__fmul_rn(cam.fx, xc), invzc), cam.cx);
  const float v = __fadd_rn(__fmul_rn(__fmul_rn(cam.fy, yc), invzc), cam.cy);
  if (isnan(u) || isnan(v)) return false;   // reference: UB; documented deviation (oracle does the same)
  if (u < cam.min_x || u > cam.max_x) return false;
  if (v < cam.min_y || v > cam.max_y) return false;
  q.g.u = u; q.g.v = v;
  q.g.r = __fmul_rn(cam.th, cam.sf[octave]);
  q.g.ur = __fsub_rn(u, __fmul_rn(cam.bf, invzc));
  if (bForward) { q.g.min_level = octave; q.g.max_level = -1; }
  else if (bBackward) { q.g.min_level = 0; q.g.max_level = octave; }
  else { q.g.min_level = octave - 1; q.g.max_level = octave + 1; }
  q.d0 = __ldg(reinterpret_cast<const uint4*>(desc));
  q.d1 = __ldg(reinterpret_cast<const uint4*>(desc) + 1);
  return true;
}

// dynamic shared memory: state[cmax] ints + taken[cmax] bytes (cmax = cur stride rounded up)
__global__ void __launch_bounds__(MATCH_THREADS) k_match_last(MatchBatch mb, MatchCam cam, int cmax) {
  extern __shared__ __align__(16) unsigned char msm[];
  __shared__ int s_off[GRID_CELLS + 1];
  __shared__ int s_cur[GRID_CELLS];
  __shared__ int ws[33];
  __shared__ int s_hist[ORBM_HISTO_LENGTH];
  __shared__ int s_keep[3];
  __shared__ int s_acc, s_pruned;
  int* state = reinterpret_cast<int*>(msm);
  uint8_t* taken = msm + (size_t)cmax * 4;
  const int p = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int nc = mb.cn[p], nl = mb.ln[p];
  const size_t co = (size_t)p * mb.cstride, lo = (size_t)p * mb.lstride;
  const float *cx = mb.cx + co, *cy = mb.cy + co, *cang = mb.cang + co, *cur = mb.curight + co;
  const int* coct = mb.coct + co;
  const uint8_t* cdesc = mb.cdesc + co * 32;
  const int* cobs = mb.cobs ? mb.cobs + co : nullptr;
  const float* lxw = mb.lxw + lo * 3;
  const uint8_t* lvalid = mb.lvalid + lo;
  const int* loct = mb.loct + lo;
  const float* lang = mb.lang + lo;
  const uint8_t* ldesc = mb.ldesc + lo * 32;
  const int* lobs = mb.lobs ? mb.lobs + lo : nullptr;
  int* grididx = mb.grididx + co;
  unsigned long long* topk = mb.topk + lo * MATCH_K;
  int* ncand = mb.ncand + lo;
  int* accepted = mb.accepted + lo;
  int* out = mb.cur2last + co;
  const float* Tc = mb.cTcw + (size_t)p * 16;
  const float* Tl = mb.lTcw + (size_t)p * 16;

  // ---- A: grid ------------------------------------------------------------------------------------------
  GridView gv;
  gv.min_x = cam.min_x; gv.min_y = cam.min_y;
  gv.inv_w = __fdiv_rn((float)GRID_COLS, __fsub_rn(cam.max_x, cam.min_x));   // src/Frame.cc:221-222
  gv.inv_h = __fdiv_rn((float)GRID_ROWS, __fsub_rn(cam.max_y, cam.min_y));
  build_grid(nc, cx, cy, gv.min_x, gv.min_y, gv.inv_w, gv.inv_h, s_off, s_cur, grididx, ws);
  gv.off = s_off; gv.idx = grididx;
  for (int j = tid; j < nc; j += nthr) {
    const bool pre = cobs && cobs[j] >= 0;
    state[j] = pre ? -2 : -1;
    taken[j] = 0;   // pre-existing entries with observations are filtered statically in walk_window
  }
  if (tid < ORBM_HISTO_LENGTH) s_hist[tid] = 0;
  if (tid == 0) { s_acc = 0; s_pruned = 0; }
  // forward / backward decision (:1590-1602)
  float twc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {   // -Rcw.t()*tcw through the generic (double-accumulating) gemm
    double s = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) s = __dadd_rn(s, __dmul_rn((double)Tc[k * 4 + i], (double)Tc[k * 4 + 3]));
    twc[i] = __double2float_rn(-s);
  }
  const float Rl2[3] = {Tl[8], Tl[9], Tl[10]};
  const float tlc2 = gemm3(Rl2, twc[0], twc[1], twc[2], Tl[11]);
  const bool bForward = (tlc2 > cam.b) && !cam.mono;
  const bool bBackward = (-tlc2 > cam.b) && !cam.mono;
  __syncthreads();

  // ---- B: independent part of every query -----------------------------------------------------------------
  for (int i = tid; i < nl; i += nthr) {
    unsigned long long k0 = ~0ull, k1 = ~0ull, k2 = ~0ull, k3 = ~0ull;
    int n_c = 0;
    if (lvalid[i]) {
      LastQuery q;
      if (setup_last_query(cam, Tc, bForward, bBackward, lxw + 3 * (size_t)i, loct[i], ldesc + 32 * (size_t)i, q)) {
        n_c = walk_window(gv, q.g, cx, cy, coct, cur, cobs, cdesc, q.d0, q.d1, true,
                          [&](int idx, int ord, int dist) {
                            const unsigned long long k = mk_key(dist, ord, idx);
                            if (k < k3) {
                              if (k < k2) {
                                k3 = k2;
                                if (k < k1) {
                                  k2 = k1;
                                  if (k < k0) { k1 = k0; k0 = k; } else k1 = k;
                                } else k2 = k;
                              } else k3 = k;
                            }
                          });
      }
    }
    topk[(size_t)i * MATCH_K + 0] = k0;
    topk[(size_t)i * MATCH_K + 1] = k1;
    topk[(size_t)i * MATCH_K + 2] = k2;
    topk[(size_t)i * MATCH_K + 3] = k3;
    ncand[i] = n_c;
    accepted[i] = -1;
  }
  __syncthreads();

  // ---- C: order-exact resolve (one warp, all lanes redundant) ---------------------------------------------
  if (tid < 32) {
    const int lane = tid;
    for (int base = 0; base < nl; base += 32) {
      const int i = base + lane;
      unsigned long long k[MATCH_K];
      int n_c = 0, obs = 0;
#pragma unroll
      for (int j = 0; j < MATCH_K; ++j) k[j] = (i < nl) ? topk[(size_t)i * MATCH_K + j] : ~0ull;
      if (i < nl) {
        n_c = ncand[i];
        obs = lobs ? lobs[i] : cam.last_obs_default;
      }
      const int lim = min(32, nl - base);
      for (int t = 0; t < lim; ++t) {
        const int qn = __shfl_sync(0xffffffffu, n_c, t);
        if (qn == 0) continue;
        const int qobs = __shfl_sync(0xffffffffu, obs, t);
        int best_idx = -1, best_dist = 256;
        bool exhausted = true;
#pragma unroll
        for (int j = 0; j < MATCH_K; ++j) {
          const unsigned long long kj = __shfl_sync(0xffffffffu, k[j], t);
          if (best_idx < 0 && kj != ~0ull) {
            const int idx = (int)(kj & 0xfffffull);
            if (!taken[idx]) { best_idx = idx; best_dist = (int)(kj >> 40); exhausted = false; }
          }
        }
        if (best_idx < 0 && qn <= MATCH_K) exhausted = false;   // every candidate is claimed: no match
        if (exhausted && best_idx < 0) {
          // more than K candidates and the K best are all claimed: re-walk the window with the claimed filter
          const int qi = base + t;
          LastQuery q;
          if (setup_last_query(cam, Tc, bForward, bBackward, lxw + 3 * (size_t)qi, loct[qi], ldesc + 32 * (size_t)qi, q)) {
            walk_window(gv, q.g, cx, cy, coct, cur, cobs, cdesc, q.d0, q.d1, true,
                        [&](int idx, int /*ord*/, int dist) {
                          if (!taken[idx] && dist < best_dist) { best_dist = dist; best_idx = idx; }
                        });
          }
        }
        if (best_idx >= 0 && best_dist <= ORBM_TH_HIGH) {
          if (lane == 0) {
            state[best_idx] = base + t;
            if (qobs > 0) taken[best_idx] = 1;
            accepted[base + t] = best_idx;
          }
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();

  // ---- D: rotation consistency (:1683-1721) -----------------------------------------------------------------
  int my_bins[4];   // up to 4 queries per thread at nl <= 4*nthr; generic loop recomputes otherwise
  (void)my_bins;
  int acc = 0;
  for (int i = tid; i < nl; i += nthr) {
    const int idx = accepted[i];
    if (idx >= 0) {
      ++acc;
      if (cam.check_ori) {
        float rot = __fsub_rn(lang[i], cang[idx]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, 1.0f / ORBM_HISTO_LENGTH));
        if (bin == ORBM_HISTO_LENGTH) bin = 0;
        atomicAdd(&s_hist[bin], 1);
      }
    }
  }
  if (acc) atomicAdd(&s_acc, acc);
  __syncthreads();
  if (cam.check_ori) {
    if (tid == 0) {   // ComputeThreeMaxima (:1912-1957)
      int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
      for (int i = 0; i < ORBM_HISTO_LENGTH; ++i) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
      s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3;
    }
    __syncthreads();
    int pr = 0;
    for (int i = tid; i < nl; i += nthr) {
      const int idx = accepted[i];
      if (idx >= 0) {
        float rot = __fsub_rn(lang[i], cang[idx]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, 1.0f / ORBM_HISTO_LENGTH));
        if (bin == ORBM_HISTO_LENGTH) bin = 0;
        if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) {
          state[idx] = -1;   // every writer stores NULL: order-free
          ++pr;
        }
      }
    }
    if (pr) atomicAdd(&s_pruned, pr);
    __syncthreads();
  }
  for (int j = tid; j < nc; j += nthr) out[j] = state[j];
  if (tid == 0) mb.nmatch[p] = s_acc - s_pruned;
}

}  // namespace b200
