// orbm_kernels.cuh -- sm_100a kernels of the ORB matcher (reference: src/ORBmatcher.cc + the Frame grid
// helpers of src/Frame.cc).  Integer / bitwise work: xor + popcount, no tensor cores.
//
// One CTA per frame pair.  Phases inside the CTA:
//   A  AssignFeaturesToGrid (src/Frame.cc:319-334): 64x48 CSR of the current frame's keypoints, cell lists in
//      ascending keypoint index (= the reference's push_back order)
//   B  per query (last-frame MapPoint), in parallel: projection, GetFeaturesInArea window walk in the
//      reference's (ix, iy, insertion) order, Hamming distances, and the TOP-K candidates by (distance, walk
//      order) -- everything that does not depend on the loop-carried "already claimed" state
//   C  order-exact resolve by one warp: queries in index order take their first unclaimed top-K entry
//      (src/ORBmatcher.cc:1656-1658 makes query i depend on the claims of queries < i); a query whose K entries
//      are all claimed re-walks its window
//   D  rotation histogram + ComputeThreeMaxima prune (:1700-1721)
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int GRID_COLS = ORBM_GRID_COLS, GRID_ROWS = ORBM_GRID_ROWS, GRID_CELLS = GRID_COLS * GRID_ROWS;
constexpr int MATCH_THREADS = 512;
constexpr int MATCH_K = 4;

struct MatchCam {            // per-call constants (Frame statics + ORBmatcher ctor args)
  float fx, fy, cx, cy, bf, b;
  float min_x, max_x, min_y, max_y;
  float sf[MAX_LEVELS];
  float th, nnratio;
  int mono, check_ori, nlevels;
  int last_obs_default;      // Observations() of last-frame MapPoints when no array is given
};

struct MatchBatch {          // SoA views; pair p reads cur arrays at p*cstride and last arrays at p*lstride
  const float *cx, *cy, *cang, *curight;
  const int* coct;
  const uint8_t* cdesc;
  const int* cobs;           // nullable
  const int* cn;
  const float* cTcw;         // p*16
  size_t cstride;
  const float* lxw;
  const uint8_t* lvalid;
  const int* loct;
  const float* lang;
  const uint8_t* ldesc;
  const int* lobs;           // nullable
  const int* ln;
  const float* lTcw;
  size_t lstride;
  int* cur2last;             // p*cstride
  int* nmatch;               // p
  unsigned long long* topk;  // p*lstride*MATCH_K
  int* ncand;                // p*lstride
  int* grididx;              // p*cstride
  int* accepted;             // p*lstride
};

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint8_t* __restrict__ b) {
  const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(b)), b1 = __ldg(reinterpret_cast<const uint4*>(b) + 1);
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// cv::Mat small-matrix gemm row: float accumulation, "+C" added in double (see oracle/match_ref.cpp header)
__device__ __forceinline__ float gemm3(const float* a, float b0, float b1, float b2, float c) {
  const float t = __fadd_rn(__fadd_rn(__fmul_rn(a[0], b0), __fmul_rn(a[1], b1)), __fmul_rn(a[2], b2));
  return __double2float_rn(__dadd_rn((double)t, (double)c));
}

struct GridView {
  const int* off;     // GRID_CELLS+1 (shared memory)
  const int* idx;     // keypoint indices, cell-major (ix*GRID_ROWS+iy), ascending inside a cell
  float min_x, min_y, inv_w, inv_h;
};

struct QueryGeom {   // what the window walk of one query needs
  float u, v, r, ur;       // projection, radius, predicted right coordinate
  int min_level, max_level;
  bool ok;
};

// GetFeaturesInArea (src/Frame.cc:465-518) fused with the candidate loop of SearchByProjection
// (src/ORBmatcher.cc:1653-1676).  Calls fn(idx, ord, dist) for every candidate that survives the static
// checks, in the reference's walk order; `ord` counts them.
template <class Fn>
__device__ __forceinline__ int walk_window(const GridView& g, const QueryGeom& q, const float* __restrict__ cx,
                                           const float* __restrict__ cy, const int* __restrict__ coct,
                                           const float* __restrict__ curight, const int* __restrict__ cobs,
                                           const uint8_t* __restrict__ cdesc, const uint4 d0, const uint4 d1,
                                           bool check_right, Fn fn) {
  const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.u, g.min_x), q.r), g.inv_w)));
  if (nMinCellX >= GRID_COLS) return 0;
  const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.u, g.min_x), q.r), g.inv_w)));
  if (nMaxCellX < 0) return 0;
  const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.v, g.min_y), q.r), g.inv_h)));
  if (nMinCellY >= GRID_ROWS) return 0;
  const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.v, g.min_y), q.r), g.inv_h)));
  if (nMaxCellY < 0) return 0;
  const bool bCheckLevels = (q.min_level > 0) || (q.max_level >= 0);
  int ord = 0;
  for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
    for (int iy = nMinCellY; iy <= nMaxCellY; ++iy) {
      const int c = ix * GRID_ROWS + iy;
      for (int e = g.off[c]; e < g.off[c + 1]; ++e) {
        const int idx = g.idx[e];
        if (bCheckLevels) {
          const int o = coct[idx];
          if (o < q.min_level) continue;
          if (q.max_level >= 0 && o > q.max_level) continue;
        }
        const float dx = __fsub_rn(cx[idx], q.u), dy = __fsub_rn(cy[idx], q.v);
        if (!(fabsf(dx) < q.r && fabsf(dy) < q.r)) continue;
        // -- from here: the candidate loop of the matcher --
        if (cobs && cobs[idx] > 0) continue;        // pre-existing MapPoint with observations: never overwritten
        if (check_right) {
          const float ur = curight[idx];
          if (ur > 0) {
            const float er = fabsf(__fsub_rn(q.ur, ur));
            if (er > q.r) continue;
          }
        }
        const int dist = hamming256(d0, d1, cdesc + (size_t)idx * 32);
        fn(idx, ord, dist);
        ++ord;
      }
    }
  }
  return ord;
}

__device__ __forceinline__ unsigned long long mk_key(int dist, int ord, int idx) {
  return ((unsigned long long)dist << 40) | ((unsigned long long)(ord & 0xfffff) << 20) | (unsigned long long)idx;
}

// Phase A.  Shared: off[GRID_CELLS+1] and cur[GRID_CELLS] ints; global: idx[n].
__device__ __forceinline__ int grid_cell(float x, float y, float min_x, float min_y, float inv_w, float inv_h) {
  // PosInGrid (src/Frame.cc:522-531): round(), not floor(); keypoints outside the 64x48 grid are dropped
  const int px = (int)roundf(__fmul_rn(__fsub_rn(x, min_x), inv_w));
  const int py = (int)roundf(__fmul_rn(__fsub_rn(y, min_y), inv_h));
  if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) return -1;
  return px * GRID_ROWS + py;
}

__device__ void build_grid(int n, const float* __restrict__ x, const float* __restrict__ y, float min_x, float min_y,
                           float inv_w, float inv_h, int* off, int* cur, int* idx, int* ws) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int c = tid; c < GRID_CELLS; c += nthr) cur[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += nthr) {
    const int c = grid_cell(x[i], y[i], min_x, min_y, inv_w, inv_h);
    if (c >= 0) atomicAdd(&cur[c], 1);
  }
  __syncthreads();
  {   // exclusive scan of the GRID_CELLS counts (contiguous chunk per thread)
    const int per = (GRID_CELLS + nthr - 1) / nthr;
    const int beg = min(GRID_CELLS, tid * per), end = min(GRID_CELLS, beg + per);
    int s = 0;
    for (int c = beg; c < end; ++c) s += cur[c];
    int total;
    int run = block_excl_scan(s, ws, &total);
    for (int c = beg; c < end; ++c) {
      off[c] = run;
      run += cur[c];
      cur[c] = 0;
    }
    if (tid == 0) off[GRID_CELLS] = total;
    __syncthreads();
  }
  for (int i = tid; i < n; i += nthr) {   // unordered fill ...
    const int c = grid_cell(x[i], y[i], min_x, min_y, inv_w, inv_h);
    if (c >= 0) idx[off[c] + atomicAdd(&cur[c], 1)] = i;
  }
  __syncthreads();
  for (int c = tid; c < GRID_CELLS; c += nthr) {   // ... then restore the push_back order (ascending index)
    const int b = off[c], e = off[c + 1];
    for (int i = b + 1; i < e; ++i) {
      const int v = idx[i];
      int j = i - 1;
      while (j >= b && idx[j] > v) { idx[j + 1] = idx[j]; --j; }
      idx[j + 1] = v;
    }
  }
  __syncthreads();
}

}  // namespace b200
