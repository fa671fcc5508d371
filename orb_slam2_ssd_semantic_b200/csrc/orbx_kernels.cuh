// orbx_kernels.cuh -- sm_100a kernels of the ORB extractor (reference: src/ORBextractor.cc).
//
// Data layout in HBM (one batch of F frames, identical geometry):
//   pyramid   level l : F planes of h_l rows x pitch_l bytes (pitch_l = w_l rounded up to 16), no border --
//             the 19-px REFLECT_101 frame of ComputePyramid (:1125-1142) is never read by anything that
//             feeds operator()'s outputs; orbx_get_level synthesises it on read-back.
//             Level 0 aliases the caller's device buffer when the batch is already resident.
//   blurred   same layout, GaussianBlur 7x7 sigma 2 of every level (:1094-1095)
//   cand      per frame: one fixed segment per FAST cell (worst-case capacity ceil(iw/2)*ceil(ih/2): strict
//             3x3 NMS forbids 8-adjacent survivors), packed u32 x:12 | y:12 | response:8, row-major in cell
//   cellcnt   per frame, per cell survivor count
//   sel       per (frame, level): quad-tree survivors in the reference's list order, same packing
//   kps/desc  per frame: cap x OrbxKeyPoint (28 B) / cap x 32 B, level-major like operator() concatenates
#pragma once
#include "common.cuh"
#define B200_HD __host__ __device__ __forceinline__
#include "../../include/glibc_sincosf.h"

namespace b200 {

struct PyrView {                 // one image pyramid (raw or blurred) of a batch
  uint8_t* p[MAX_LEVELS];        // frame 0 of level l
  size_t fstride[MAX_LEVELS];    // bytes between consecutive frames of level l
  int pitch[MAX_LEVELS];
  int w[MAX_LEVELS], h[MAX_LEVELS];
};

struct CellDesc {                // one FAST cell (src/ORBextractor.cc:798-838)
  short level;
  short x0, y0;                  // ROI origin in level coordinates (= iniX, iniY)
  short rw, rh;                  // ROI size (maxX-iniX, maxY-iniY); detection range is the ROI minus 3 px
  short pad;
  int slot_off;                  // offset of this cell's segment inside a frame's cand[]
};

struct LevelTab {                // per-level constants of the extractor
  int cell_begin[MAX_LEVELS + 1];   // cells of level l: [cell_begin[l], cell_begin[l+1])
  int slot_begin[MAX_LEVELS + 1];   // cand slots of level l
  int nfeat[MAX_LEVELS];            // mnFeaturesPerLevel
  int sel_off[MAX_LEVELS + 1];      // offset of level l inside a frame's sel[] / capacity
  int n_ini[MAX_LEVELS];            // DistributeOctTree nIni (:545)
  int box_h[MAX_LEVELS];            // maxBorderY-minBorderY: height of the initial nodes (:558)
  float hx[MAX_LEVELS];             // DistributeOctTree hX (:547)
  float sf[MAX_LEVELS];             // mvScaleFactor
  float kp_size[MAX_LEVELS];        // (float)(int)(PATCH_SIZE*mvScaleFactor[l]) (:846)
  int nlevels;
};

__device__ __forceinline__ unsigned pack_kp(int x, int y, int resp) {
  return (unsigned)x | ((unsigned)y << 12) | ((unsigned)resp << 24);
}
__device__ __forceinline__ int kp_x(unsigned v) { return v & 0xfff; }
__device__ __forceinline__ int kp_y(unsigned v) { return (v >> 12) & 0xfff; }
__device__ __forceinline__ int kp_r(unsigned v) { return v >> 24; }

// ---------------------------------------------------------------------------------------------------
// K1  pyramid level l from level l-1: cv::resize INTER_LINEAR 8UC1 (call site :1134).
//     xt/yt: per destination index {source offset, a0 | a1<<16} with the 11-bit coefficients.
//     One thread -> 4 horizontally adjacent destination pixels -> one 32-bit store.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize(const uint8_t* __restrict__ src, int spitch, size_t sfs, int sw,
                                                int sh, uint8_t* __restrict__ dst, int dpitch, size_t dfs,
                                                int dw, int dh, const int2* __restrict__ xt,
                                                const int2* __restrict__ yt) {
  const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int dy = blockIdx.y * 8 + threadIdx.y;
  if (dx0 >= dw || dy >= dh) return;
  const uint8_t* s = src + (size_t)blockIdx.z * sfs;
  uint8_t* d = dst + (size_t)blockIdx.z * dfs;
  const int2 ty = __ldg(&yt[dy]);
  const int sy0 = ty.x, sy1 = min(sy0 + 1, sh - 1);
  const int b0 = (short)(ty.y & 0xffff), b1 = (short)(ty.y >> 16);
  const uint8_t* r0 = s + (size_t)sy0 * spitch;
  const uint8_t* r1 = s + (size_t)sy1 * spitch;
  unsigned out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int dx = dx0 + i;
    if (dx < dw) {
      const int2 tx = __ldg(&xt[dx]);
      const int x0 = tx.x, x1 = min(x0 + 1, sw - 1);
      const int a0 = (short)(tx.y & 0xffff), a1 = (short)(tx.y >> 16);
      const int h0 = r0[x0] * a0 + r0[x1] * a1;
      const int h1 = r1[x0] * a0 + r1[x1] * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      out |= (unsigned)(v & 0xff) << (8 * i);
    }
  }
  *reinterpret_cast<unsigned*>(d + (size_t)dy * dpitch + dx0) = out;   // pitch is a multiple of 16
}

// K1 (group-table path): a thread's 4 adjacent outputs start at a fixed destination column (a multiple of 4), so
// everything that depends only on the columns is tabulated per GROUP: first source word, byte phase, the four PRMT
// selectors (relative to the group's first source byte) and the four coefficient pairs -- two 128-bit loads.  The two
// source rows are funnel-shifted into 8-byte windows that start at the group's first source byte; each output then
// costs one PRMT + one IDP.2A per row plus the vertical pass.  Same integer arithmetic as k_resize (needs scale <= 2:
// the 4 outputs read at most 8 consecutive source bytes).
__global__ void __launch_bounds__(256) k_resize_g(const uint8_t* __restrict__ src, int spitch, size_t sfs, int sh,
                                                  uint8_t* __restrict__ dst, int dpitch, size_t dfs, int dw, int dh,
                                                  const int4* __restrict__ xg /* 2 per group */,
                                                  const int2* __restrict__ yt) {
  const int g = blockIdx.x * 32 + threadIdx.x;
  const int dx0 = g * 4;
  const int dy = blockIdx.y * 8 + threadIdx.y;
  if (dx0 >= dw || dy >= dh) return;
  const uint8_t* s = src + (size_t)blockIdx.z * sfs;
  uint8_t* d = dst + (size_t)blockIdx.z * dfs;
  const int2 ty = __ldg(&yt[dy]);
  const int sy0 = ty.x, sy1 = min(sy0 + 1, sh - 1);
  const int b0 = (short)(ty.y & 0xffff), b1 = (short)(ty.y >> 16);
  const int4 ga = __ldg(&xg[2 * g]), gc = __ldg(&xg[2 * g + 1]);   // {word, 8*phase, selectors, -} / coefficients
  const int wlast = (spitch >> 2) - 1;                              // never read past the row's last word
  const int w0 = ga.x, w1 = min(ga.x + 1, wlast), w2 = min(ga.x + 2, wlast);
  const unsigned* r0 = reinterpret_cast<const unsigned*>(s + (size_t)sy0 * spitch);
  const unsigned* r1 = reinterpret_cast<const unsigned*>(s + (size_t)sy1 * spitch);
  const unsigned a0 = __ldg(r0 + w0), a1 = __ldg(r0 + w1), a2 = __ldg(r0 + w2);
  const unsigned c0 = __ldg(r1 + w0), c1 = __ldg(r1 + w1), c2 = __ldg(r1 + w2);
  const unsigned ph = (unsigned)ga.y;
  const unsigned A_lo = __funnelshift_r(a0, a1, ph), A_hi = __funnelshift_r(a1, a2, ph);   // bytes sx0 .. sx0+7
  const unsigned C_lo = __funnelshift_r(c0, c1, ph), C_hi = __funnelshift_r(c1, c2, ph);
  const unsigned sels = (unsigned)ga.z;
  const int co[4] = {gc.x, gc.y, gc.z, gc.w};
  unsigned out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = sels >> (8 * i);   // low two nibbles: bytes d_i, d_i+1 of the window
    const unsigned p0 = __byte_perm(A_lo, A_hi, sel), p1 = __byte_perm(C_lo, C_hi, sel);
    const int h0 = (int)__dp2a_lo((unsigned)co[i], p0, 0u);   // a0*s[x] + a1*s[x+1]
    const int h1 = (int)__dp2a_lo((unsigned)co[i], p1, 0u);
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    out |= (unsigned)(v & 0xff) << (8 * i);
  }
  *reinterpret_cast<unsigned*>(d + (size_t)dy * dpitch + dx0) = out;   // pitch is a multiple of 16
}

// ---------------------------------------------------------------------------------------------------
// K2  per-cell FAST-9/16 + strict 3x3 NMS with the ini/min threshold fallback
//     (cv::FAST TYPE_9_16 nonmax=true; call sites :818,:823; cell loop :798-838).
//     One CTA per (cell, frame).  m(p) = max over 16 arcs of max(min d, min -d) is threshold-independent,
//     so it is evaluated once per pixel; "corner at t" is m > t and the score is m-1.  Survivors are
//     written row-major into the cell's fixed segment, so concatenating the segments cell-major reproduces
//     vToDistributeKeys' order.
// ---------------------------------------------------------------------------------------------------
constexpr int FAST_WARPS = 8;
constexpr int FAST_THREADS = FAST_WARPS * 32;
constexpr int FAST_MAX_ROI = 72;   // ROI side bound enforced at create (cell <= 60 px + 6, padded)
constexpr int FAST_TP_SMALL = 48;  // compile-time tile pitches (>= 3 + ROI width, multiple of 4): ring offsets are
constexpr int FAST_TP_BIG = 80;
constexpr int FAST_QLEN = 256;     // per-warp ring of pixels that passed the compass test (<= 31 pending + 128 new)    // immediates; SMALL serves ROIs up to 45 px (640x480: 43), BIG the general case

// m(p) of one pixel.  Packing: one IMAD per ring pixel gives lo16 = 256 + (c - r), hi16 = 256 + (r - c) (biased,
// both in [1, 511], so no borrow crosses the halves); min3/max3 on s16x2 then evaluate the bright and the dark
// arcs at once: a3[i] = min(v[i..i+2]), a9[i] = min(a3[i], a3[i+3], a3[i+6]) = min over the 9-arc starting at i.
// P = row pitch of the tile in bytes.
template <int P>
__device__ __forceinline__ int fast_m_exact(const uint8_t* c) {
  const int bias = 256 * 65537 - (int)c[0] * 65535;
  unsigned v[16];
  v[0] = (unsigned)((int)c[3 * P] * 65535 + bias);       v[1] = (unsigned)((int)c[3 * P + 1] * 65535 + bias);
  v[2] = (unsigned)((int)c[2 * P + 2] * 65535 + bias);   v[3] = (unsigned)((int)c[P + 3] * 65535 + bias);
  v[4] = (unsigned)((int)c[3] * 65535 + bias);           v[5] = (unsigned)((int)c[-P + 3] * 65535 + bias);
  v[6] = (unsigned)((int)c[-2 * P + 2] * 65535 + bias);  v[7] = (unsigned)((int)c[-3 * P + 1] * 65535 + bias);
  v[8] = (unsigned)((int)c[-3 * P] * 65535 + bias);      v[9] = (unsigned)((int)c[-3 * P - 1] * 65535 + bias);
  v[10] = (unsigned)((int)c[-2 * P - 2] * 65535 + bias); v[11] = (unsigned)((int)c[-P - 3] * 65535 + bias);
  v[12] = (unsigned)((int)c[-3] * 65535 + bias);         v[13] = (unsigned)((int)c[P - 3] * 65535 + bias);
  v[14] = (unsigned)((int)c[2 * P - 2] * 65535 + bias);  v[15] = (unsigned)((int)c[3 * P - 1] * 65535 + bias);
  unsigned a3[16], a9[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a3[i] = __vimin3_s16x2(v[i], v[(i + 1) & 15], v[(i + 2) & 15]);
#pragma unroll
  for (int i = 0; i < 16; ++i) a9[i] = __vimin3_s16x2(a3[i], a3[(i + 3) & 15], a3[(i + 6) & 15]);
  unsigned m0 = __vimax3_s16x2(a9[0], a9[1], a9[2]);
  unsigned m1 = __vimax3_s16x2(a9[3], a9[4], a9[5]);
  unsigned m2 = __vimax3_s16x2(a9[6], a9[7], a9[8]);
  unsigned m3 = __vimax3_s16x2(a9[9], a9[10], a9[11]);
  unsigned m4 = __vimax3_s16x2(a9[12], a9[13], a9[14]);
  m0 = __vimax3_s16x2(m0, m1, m2);
  m3 = __vimax3_s16x2(m3, m4, a9[15]);
  m0 = __vmaxs2(m0, m3);
  // halves hold 256 + min(d) over the best bright arc / 256 + min(-d) over the best dark arc
  return max(0, max((int)(m0 & 0xffffu), (int)(m0 >> 16)) - 256);
}

// One WARP per (cell, frame), 8 cells per CTA, no block-level synchronisation.  Per-warp shared memory: the ROI and
// its score map interleaved row by row (pixel row y at y*2*TP, its score row at y*2*TP + TP, so one base register and
// compile-time offsets address both), a 64-entry ring of pixels that passed the compass test, and the list of corners.
//   sweep   lane = column, row by row (a second sweep covers cells wider than 32): compass test; failing pixels get
//           score 0, passing ones are queued;
//   dense   whenever 32 pixels are queued, a full warp evaluates their exact m(p); corners (m > t) go to the corner
//           list -- the queue is FIFO over a row-major sweep, so the list is row-major as well;
//   NMS     32 corners at a time: strict 3x3 maximum, order-preserving compaction into the cell's output segment.
// `aligned` bit 0 (host-checked): level rows are 4-byte aligned, so the ROI is fetched as aligned 32-bit words and kept
// at the same byte phase (pixel x of the ROI sits at column (x0 & 3) + x).  Bit 1 selects the second tile staging:
// up to four ROI rows per warp step and 128-bit stores for the score-map clear (the first one spent 8 % of the
// kernel's instructions there, profiles/r02_ncu_v4_summary.txt).
template <int TP, bool SWEEP4>
__global__ void __launch_bounds__(FAST_THREADS, 4) k_fast_cells(PyrView pyr, const CellDesc* __restrict__ cells,
                                                                int ncells, int slots_per_frame, int ini_th,
                                                                int min_th, unsigned* __restrict__ cand,
                                                                int* __restrict__ cellcnt, int aligned, int rows_max,
                                                                int clist_cap) {
  extern __shared__ __align__(16) uint8_t fsm[];
  constexpr int P = 2 * TP;   // row pitch of the interleaved tile
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cell = blockIdx.x * (blockDim.x >> 5) + w, f = blockIdx.y;   // blockDim.x / 32 cells per CTA (host: <= FAST_WARPS)
  if (cell >= ncells) return;
  uint8_t* tile = fsm + (size_t)w * (rows_max * P + 2 * FAST_QLEN + 2 * clist_cap);
  unsigned short* queue = reinterpret_cast<unsigned short*>(tile + (size_t)rows_max * P);
  unsigned short* clist = queue + FAST_QLEN;   // corners at t, row-major
  const CellDesc cd = cells[cell];
  const int l = cd.level, rw = cd.rw, rh = cd.rh;
  const int pitch = pyr.pitch[l];
  const bool stage2 = (aligned & 2) != 0;     // host switch: second formulation of the tile staging (below)
  aligned &= 1;
  const int sh = aligned ? (cd.x0 & 3) : 0;   // byte phase of the ROI inside its first word
  {
    const uint8_t* img = pyr.p[l] + (size_t)f * pyr.fstride[l] + (size_t)cd.y0 * pitch + (cd.x0 - sh);
    if (aligned) {
      const int nw = (sh + rw + 3) >> 2;   // words per ROI row (<= (3 + FAST_MAX_ROI + 3) / 4 = 19)
      if (stage2) {
        // up to four rows per warp step: lane = r * nw + k (640x480: nw = 10 -> 3 rows, 30 lanes); no integer division
        const int rpl = (nw <= 8) ? 4 : (nw <= 10) ? 3 : (nw <= 16) ? 2 : 1;
        const int r0 = (lane >= nw) + (lane >= 2 * nw) + (lane >= 3 * nw), k = lane - r0 * nw;
        if (r0 < rpl && k < nw) {
          const unsigned* src = reinterpret_cast<const unsigned*>(img + (size_t)r0 * pitch) + k;
          unsigned* dst = reinterpret_cast<unsigned*>(tile + r0 * P) + k;
          const size_t sstep = (size_t)rpl * (pitch >> 2);   // level pitches are multiples of 4 on this path
          for (int y = r0; y < rh; y += rpl, src += sstep, dst += rpl * (P / 4)) *dst = __ldg(src);
        }
      } else if (nw <= 16) {               // two rows per step: lanes 0-15 / 16-31
        const int k = lane & 15, half = lane >> 4;
        for (int y = half; y < rh; y += 2)
          if (k < nw) reinterpret_cast<unsigned*>(tile + y * P)[k] = __ldg(reinterpret_cast<const unsigned*>(img + (size_t)y * pitch) + k);
      } else {
        for (int y = 0; y < rh; ++y)
          for (int k = lane; k < nw; k += 32)
            reinterpret_cast<unsigned*>(tile + y * P)[k] = __ldg(reinterpret_cast<const unsigned*>(img + (size_t)y * pitch) + k);
      }
    } else {
      for (int y = 0; y < rh; ++y)
        for (int x = lane; x < rw; x += 32) tile[y * P + x] = __ldg(img + (size_t)y * pitch + x);
    }
    if (SWEEP4 && aligned) {
      // the word sweep only writes the scores of corners: clear every score row once (rows 2 .. rh-3 are read by the NMS)
      if (stage2) {   // tile, P and TP are multiples of 16 bytes: 128-bit stores
        constexpr int QPR = TP / 16;
        for (int i = lane; i < (rh - 4) * QPR; i += 32) {
          const int y = 2 + i / QPR, k = i - (y - 2) * QPR;
          reinterpret_cast<uint4*>(tile + y * P + TP)[k] = make_uint4(0u, 0u, 0u, 0u);
        }
      } else {
        constexpr int WPR = TP / 4;
        for (int i = lane; i < (rh - 4) * WPR; i += 32) {
          const int y = 2 + i / WPR, k = i - (y - 2) * WPR;
          reinterpret_cast<unsigned*>(tile + y * P + TP)[k] = 0u;
        }
      }
    } else {
      // only the one-pixel frame around the detection range is read without being written: clear it
      for (int i = lane; i < rw; i += 32) { tile[2 * P + TP + sh + i] = 0; tile[(rh - 3) * P + TP + sh + i] = 0; }
      for (int i = lane; i < rh; i += 32) { tile[i * P + TP + sh + 2] = 0; tile[i * P + TP + sh + rw - 3] = 0; }
    }
  }
  __syncwarp();
  unsigned* out = cand + (size_t)f * slots_per_frame + cd.slot_off;
  const bool wide = rw - 6 > 32;
  const unsigned lt_mask = (1u << lane) - 1u;
  int total = 0;
  // pass 0: everything at ini_th (pixels with m <= ini_th can neither be corners nor outscore one at that threshold);
  // pass 1 (:821, only when the cell is EMPTY AFTER non-max suppression): the same at min_th.
  for (int pass = 0; pass < 2 && total == 0; ++pass) {
    const int t = pass ? min_th : ini_th;
    if (pass == 1 && ini_th == min_th) break;
    int qh = 0, qn = 0;   // candidate ring (warp-uniform head / fill)
    int cn = 0;           // corners at t found so far (warp-uniform)
    // exact m of the queued pixel `o` (tile offset) of the lanes with `on`, score map + corner list update
    auto dense = [&](int o, bool on) {
      bool cr = false;
      if (on) {
        const int m = fast_m_exact<P>(tile + o);
        cr = m > t;
        tile[o + TP] = (uint8_t)(cr ? m : 0);
      }
      const unsigned cb = __ballot_sync(0xffffffffu, cr);
      if (cr) clist[cn + __popc(cb & lt_mask)] = (unsigned short)o;
      cn += __popc(cb);
    };
    if (SWEEP4 && aligned) {
      // Word sweep: a thread owns the 4 pixels of one aligned tile word (columns 4g .. 4g+3) of one row; several rows per
      // warp step.  The compass test runs on pixel PAIRS in s16x2 lanes, bright and dark polarity in separate registers:
      //   bright: 256 + r - c,  dark: 256 + c - r   (one IADD3 each per ring pixel and pixel pair, operands spread from the
      //   byte lanes by one PRMT), q = min(max(v0, v8), max(v4, v12)) per polarity, hit <=> q > 256 + t in either.
      // Queue positions are the pixels' ranks in ROW-MAJOR order (lanes are ordered (row, word); inside a word by column),
      // so the corner list the dense phase builds stays row-major.
      const unsigned kq = (unsigned)(0x7fff - 256 - t) * 0x10001u;
      const int gA = (sh + 3) >> 2, gB = (sh + rw - 4) >> 2, NW = gB - gA + 1;
      const int RPI = 32 / NW;                       // rows per warp step (NW <= 19 < 32)
      const int rs = lane / NW, g = gA + (lane - rs * NW);
      const int cmin = sh + 3, cmax = sh + rw - 3;   // detection columns [cmin, cmax) in tile coordinates
      const unsigned* trow = reinterpret_cast<const unsigned*>(tile);
      for (int y0 = 3; y0 < rh - 3; y0 += RPI) {
        const int y = y0 + rs;
        unsigned hits = 0;                           // bit k: pixel 4g+k of row y passed
        if (rs < RPI && y < rh - 3) {
          const unsigned* wr = trow + (y * P) / 4;
          const unsigned wc = wr[g], wl = wr[max(g - 1, 0)], wn = wr[g + 1];
          const unsigned w0 = wr[g + (3 * P) / 4], w8 = wr[g - (3 * P) / 4];
          const unsigned w12 = __funnelshift_r(wl, wc, 8);     // bytes x-3 of the four pixels
          const unsigned w4 = __funnelshift_r(wc, wn, 24);     // bytes x+3
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {                     // pixel pairs (0,2) and (1,3)
            const unsigned sel = h2 ? 0x4341u : 0x4240u;
            const unsigned C = __byte_perm(wc, 0u, sel), R0 = __byte_perm(w0, 0u, sel), R8 = __byte_perm(w8, 0u, sel);
            const unsigned R4 = __byte_perm(w4, 0u, sel), R12 = __byte_perm(w12, 0u, sel);
            const unsigned cb = 0x01000100u - C, cd = 0x01000100u + C;   // halves stay in [1, 511]: no borrow / carry across
            const unsigned qb = __vmins2(__vmaxs2(R0 + cb, R8 + cb), __vmaxs2(R4 + cb, R12 + cb));
            const unsigned qd = __vmins2(__vmaxs2(cd - R0, cd - R8), __vmaxs2(cd - R4, cd - R12));
            const unsigned hb = ((qb + kq) | (qd + kq)) & 0x80008000u;
            if (hb & 0x8000u) hits |= 1u << h2;
            if (hb & 0x80000000u) hits |= 4u << h2;
          }
          const int c0 = 4 * g;                                // drop the pixels outside the detection columns
          unsigned cm = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) cm |= ((c0 + k >= cmin) && (c0 + k < cmax)) ? (1u << k) : 0u;
          hits &= cm;
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, hits & 1u), b1 = __ballot_sync(0xffffffffu, hits & 2u);
        const unsigned b2 = __ballot_sync(0xffffffffu, hits & 4u), b3 = __ballot_sync(0xffffffffu, hits & 8u);
        if (hits) {
          int pos = qh + qn + __popc(b0 & lt_mask) + __popc(b1 & lt_mask) + __popc(b2 & lt_mask) + __popc(b3 & lt_mask);
          const int o = y * P + 4 * g;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (hits & (1u << k)) { queue[pos & (FAST_QLEN - 1)] = (unsigned short)(o + k); ++pos; }
        }
        qn += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
        while (qn >= 32) {
          __syncwarp();
          dense(queue[(qh + lane) & (FAST_QLEN - 1)], true);
          qh += 32;
          qn -= 32;
        }
      }
      __syncwarp();
      dense((lane < qn) ? queue[(qh + lane) & (FAST_QLEN - 1)] : 0, lane < qn);
    } else {
      // The compass test in packed form: a 9-arc always holds one pixel of each opposite compass pair, and all its
      // pixels lie on the same side of the centre, so min(max(v0, v8), max(v4, v12)) > 256 + t in either half is
      // necessary for a corner at t.
      const unsigned kq = (unsigned)(0x7fff - 256 - t) * 0x10001u;
      uint8_t* rowp = tile + 3 * P + sh + 3 + lane;
      const int colmax = rw - 6;   // detection columns
      for (int y = 3; y < rh - 3; ++y, rowp += P) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !wide) break;
          uint8_t* cc = rowp + 32 * h;
          bool hit = false;
          if (lane + 32 * h < colmax) {
            const int bias = 256 * 65537 - (int)cc[0] * 65535;
            const unsigned v0 = (unsigned)((int)cc[3 * P] * 65535 + bias), v8 = (unsigned)((int)cc[-3 * P] * 65535 + bias);
            const unsigned v4 = (unsigned)((int)cc[3] * 65535 + bias), v12 = (unsigned)((int)cc[-3] * 65535 + bias);
            const unsigned qv = __vmins2(__vmaxs2(v0, v8), __vmaxs2(v4, v12));
            hit = ((qv + kq) & 0x80008000u) != 0u;   // halves are in [1,511]: no carry between them
            if (!hit) cc[TP] = 0;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, hit);
          if (hit) queue[(qh + qn + __popc(bal & lt_mask)) & (FAST_QLEN - 1)] = (unsigned short)(cc - tile);
          qn += __popc(bal);
          if (qn >= 32) {
            __syncwarp();
            dense(queue[(qh + lane) & (FAST_QLEN - 1)], true);
            qh += 32;
            qn -= 32;
          }
        }
      }
      __syncwarp();
      dense((lane < qn) ? queue[(qh + lane) & (FAST_QLEN - 1)] : 0, lane < qn);
    }
    __syncwarp();
    // NMS (strict 3x3 maximum of the scores; a neighbour that is no corner at t scores 0) over the corner list, 32
    // corners at a time, + order-preserving compaction (one ballot per batch, running offset).
    for (int base = 0; base < cn; base += 32) {
      bool keep = false;
      int m = 0, o = 0;
      if (base + lane < cn) {
        o = clist[base + lane];
        const uint8_t* qq = tile + o + TP;
        m = qq[0];
        const int n0 = max(max((int)qq[-P - 1], (int)qq[-P]), (int)qq[-P + 1]);
        const int n1 = max(max((int)qq[-1], (int)qq[1]), (int)qq[P - 1]);
        const int n2 = max((int)qq[P], (int)qq[P + 1]);
        keep = m > max(max(max(n0, n1), n2), 1);   // score m-1 vs neighbour scores (m_q-1 if m_q > t, else 0)
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int y = o / P, x = o - y * P - sh;
        out[total + __popc(bal & lt_mask)] = pack_kp(cd.x0 + x, cd.y0 + y, m - 1);
      }
      total += __popc(bal);
    }
    __syncwarp();
  }
  if (lane == 0) cellcnt[(size_t)f * ncells + cell] = total;
}

// ---------------------------------------------------------------------------------------------------
// K3  DistributeOctTree (src/ORBextractor.cc:540-765) + DivideNode (:478-534), one CTA per (level, frame).
//     The std::list is held as an array in list order (front = index 0); one "round" splits a set of nodes
//     in a given processing order and rebuilds the list exactly as the push_front/erase sequence would:
//       [children of the LAST processed node (n4,n3,n2,n1) ... children of the FIRST processed node]
//       ++ [untouched nodes in their old relative order].
//     Sweep rounds (:608-667) process every node holding >1 keypoints in list order.  Largest-first rounds
//     (:678-739) process them by (size desc, creation desc); nodes created in one round sit in the list in
//     reverse creation order, so "later created first" == "smaller list index first" -- no addresses needed
//     (the reference's heap-address tie-break is replaced by this documented surrogate).
//     Keypoints keep their original (cell-major, row-major) index, so "first key with maximal response"
//     (:752-759) is an atomicMax over (response << 24 | ~index).
// ---------------------------------------------------------------------------------------------------
constexpr int QT_THREADS = 512;
constexpr int QT_UNTOUCHED = 0x40000000;   // flag in QtSmem::cpos: the node keeps its box this round

struct QtScratchView {
  unsigned* qkp;    // [F][slots_per_frame] gathered candidates (contiguous per level at slot_begin[l])
  int* qnode;       // [F][slots_per_frame] list index of each candidate's node
};

struct QtSmem {   // carved from dynamic shared memory, cap entries each
  short4* box[2];
  int* cnt[2];
  int* cc;        // [cap][4] child counts (per quadrant) of every node of the current list
  int* cc2;       // [cap][4] the same for the list being built (filled while the keypoints are re-homed)
  int* cpos;      // [cap][4] new list index of each child
  int* rk;        // processing rank or -1
  int* npos;      // new list index of untouched nodes
  int* byrank;    // node index by processing rank
  int* pre;       // inclusive prefix of child counts by rank
};

__host__ __device__ inline size_t qt_smem_bytes(int cap) { return (size_t)cap * (2 * 8 + 2 * 4 + 16 + 16 + 16 + 4 * 4); }

__device__ __forceinline__ int qt_quadrant(short4 b, unsigned kp) {
  // DivideNode's assignment (:509-523); node boxes are relative to (minBorderX, minBorderY) = (16,16)
  const int x = kp_x(kp) - FAST_BORDER, y = kp_y(kp) - FAST_BORDER;
  const int mx = b.x + ((b.z - b.x + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)
  const int my = b.y + ((b.w - b.y + 1) >> 1);
  return (x < mx) ? ((y < my) ? 0 : 2) : ((y < my) ? 1 : 3);
}

// Shared-memory counter increment with warp aggregation: the early rounds of the quad-tree funnel tens of thousands of
// keypoints into a handful of counters (1 initial node, 4, 16, ... quadrant counters), where plain atomicAdd serialises
// 32-way inside every warp.  Lanes that hit the same counter elect a leader that adds the group's population.
// All 32 lanes must call; `active` masks the tail.
__device__ __forceinline__ void qt_count(int* ctr, int idx, bool active) {
  const unsigned act = __ballot_sync(0xffffffffu, active);
  if (active) {
    const unsigned grp = __match_any_sync(act, idx);
    if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&ctr[idx], __popc(grp));
  }
}

// in-place exclusive scan of a[0..n) (shared memory); returns the total. All threads must call.
__device__ int qt_scan_array(int* a, int n, int* ws) {
  const int per = (n + QT_THREADS - 1) / QT_THREADS;   // launch contract: blockDim.x == QT_THREADS (a shift, not a division)
  const int beg = min(n, (int)threadIdx.x * per), end = min(n, beg + per);
  int s = 0;
  for (int i = beg; i < end; ++i) s += a[i];
  int total;
  int run = block_excl_scan(s, ws, &total);
  for (int i = beg; i < end; ++i) {
    const int v = a[i];
    a[i] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

// The body is shared by three entry points that differ only in their register budget (k_quadtree: the compiler's choice,
// 53 registers -> 2 CTAs per SM; k_quadtree_o3 / _o4: capped for 3 / 4 CTAs per SM): the kernel is latency-bound
// (42 % issue-active, barrier + L2 stalls, profiles/r02_ncu_v4_summary.txt), so resident CTAs are what hides it.
__device__ __forceinline__ void quadtree_body(const LevelTab& lt, const CellDesc* __restrict__ cells,
                                              const unsigned* __restrict__ cand,
                                              const int* __restrict__ cellcnt, int ncells,
                                              int slots_per_frame, QtScratchView sc, int qt_cap,
                                              unsigned* __restrict__ sel, int* __restrict__ selcnt,
                                              int* __restrict__ candcnt, int sel_per_frame,
                                              int level_begin, int kp_smem_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int ws[33];
  __shared__ int s_coff[QT_THREADS];
  __shared__ int s_expand, s_rstar;
  const int l = blockIdx.x + level_begin, f = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
  const int nlev = lt.nlevels;
  const int N = lt.nfeat[l];
  QtSmem S;
  {
    unsigned char* p = smem_raw;
    S.box[0] = (short4*)p; p += (size_t)qt_cap * 8;
    S.box[1] = (short4*)p; p += (size_t)qt_cap * 8;
    S.cnt[0] = (int*)p; p += (size_t)qt_cap * 4;
    S.cnt[1] = (int*)p; p += (size_t)qt_cap * 4;
    S.cc = (int*)p; p += (size_t)qt_cap * 16;
    S.cc2 = (int*)p; p += (size_t)qt_cap * 16;
    S.cpos = (int*)p; p += (size_t)qt_cap * 16;
    S.rk = (int*)p; p += (size_t)qt_cap * 4;
    S.npos = (int*)p; p += (size_t)qt_cap * 4;
    S.byrank = (int*)p; p += (size_t)qt_cap * 4;
    S.pre = (int*)p;
  }
  const unsigned* fcand = cand + (size_t)f * slots_per_frame;
  const int* fcnt = cellcnt + (size_t)f * ncells;
  // Per-candidate arrays (packed keypoint + node id) stay in the L2-resident global scratch: staging them in shared
  // memory was measured slower on B200 (profiles/r01_notes.md) -- it costs CTA-level concurrency and generic addressing.
  unsigned* __restrict__ qkp = sc.qkp + (size_t)f * slots_per_frame + lt.slot_begin[l];
  int* __restrict__ qnode = sc.qnode + (size_t)f * slots_per_frame + lt.slot_begin[l];
  (void)kp_smem_cap;
  // ---- gather this level's per-cell segments into one contiguous, order-preserving list -----------
  const int cb = lt.cell_begin[l], nc = lt.cell_begin[l + 1] - cb;
  int K = 0;
  for (int base = 0; base < nc; base += nthr) {
    const int c = base + tid;
    const int cn = (c < nc) ? fcnt[cb + c] : 0;
    int tot;
    const int off = block_excl_scan(cn, ws, &tot);
    s_coff[tid] = K + off;
    __syncthreads();
    const int nw = nthr >> 5, w = tid >> 5, lane = tid & 31;
    const int m = min(nthr, nc - base);
    for (int ci = w; ci < m; ci += nw) {
      const int n_c = fcnt[cb + base + ci];
      const unsigned* srcp = fcand + cells[cb + base + ci].slot_off;
      unsigned* dstp = qkp + s_coff[ci];
      for (int i = lane; i < n_c; i += 32) dstp[i] = srcp[i];
    }
    K += tot;
    __syncthreads();
  }
  if (tid == 0) candcnt[(size_t)f * nlev + l] = K;
  if (K == 0) {   // DistributeOctTree on an empty input returns an empty list
    if (tid == 0) selcnt[(size_t)f * nlev + l] = 0;
    return;
  }

  // ---- initial nodes (:553-587) ---------------------------------------------------------------------
  const int nIni = lt.n_ini[l];
  const float hX = lt.hx[l];
  const int boxH = lt.box_h[l];
  int cur = 0;
  int* icnt = S.cc;   // nIni counters
  for (int i = tid; i < nIni; i += nthr) icnt[i] = 0;
  __syncthreads();
  for (int k = tid; k < K; k += nthr) {
    const float xr = (float)(kp_x(qkp[k]) - FAST_BORDER);
    int ii = (int)__fdiv_rn(xr, hX);
    ii = min(ii, nIni - 1);
    qnode[k] = ii;
    atomicAdd(&icnt[ii], 1);
  }
  __syncthreads();
  for (int i = tid; i < nIni; i += nthr) S.npos[i] = (icnt[i] > 0) ? 1 : 0;
  __syncthreads();
  int n = qt_scan_array(S.npos, nIni, ws);
  for (int i = tid; i < nIni; i += nthr) {
    if (icnt[i] > 0) {
      const int p = S.npos[i];
      S.box[0][p] = make_short4((short)(int)__fmul_rn(hX, (float)i), 0, (short)(int)__fmul_rn(hX, (float)(i + 1)),
                                (short)boxH);
      S.cnt[0][p] = icnt[i];
    }
  }
  for (int i = tid; i < nIni * 4; i += nthr) S.cc2[i] = 0;
  __syncthreads();
  // compact the node ids and count, per node, the keypoints of each quadrant (the child populations of round 0)
  for (int k = tid; k < K; k += nthr) {
    const int p = S.npos[qnode[k]];
    qnode[k] = p;
    atomicAdd(&S.cc2[p * 4 + qt_quadrant(S.box[0][p], qkp[k])], 1);
  }
  __syncthreads();
  { int* t = S.cc; S.cc = S.cc2; S.cc2 = t; }   // S.cc aliased icnt until here

  // ---- rounds -------------------------------------------------------------------------------------------
  bool largest = false, finish = false;
  while (!finish) {
    const int prevSize = n;
    short4* box = S.box[cur];
    int* cnt = S.cnt[cur];
    short4* nbox = S.box[cur ^ 1];
    int* ncnt = S.cnt[cur ^ 1];
    // (1) processing rank of every node holding more than one keypoint
    int m;   // number of candidates
    if (!largest) {
      for (int i = tid; i < n; i += nthr) S.rk[i] = (cnt[i] > 1) ? 1 : 0;
      __syncthreads();
      m = qt_scan_array(S.rk, n, ws);
      for (int i = tid; i < n; i += nthr) {
        if (cnt[i] > 1) S.byrank[S.rk[i]] = i;
        else S.rk[i] = -1;
      }
    } else {
      for (int i = tid; i < n; i += nthr) S.npos[i] = (cnt[i] > 1) ? 1 : 0;
      __syncthreads();
      m = qt_scan_array(S.npos, n, ws);   // only the total is needed
      for (int i = tid; i < n; i += nthr) {
        const int ci = cnt[i];
        if (ci > 1) {
          int r = 0;
          for (int j = 0; j < n; ++j) {
            const int cj = cnt[j];
            r += (cj > ci) || (cj == ci && j < i);   // (size desc, list index asc); cj>ci>1 implies candidate
          }
          S.rk[i] = r;
          S.byrank[r] = i;
        } else {
          S.rk[i] = -1;
        }
      }
    }
    __syncthreads();
    // (2) the child populations S.cc[node][quadrant] were counted while the keypoints were re-homed last round
    // (3) inclusive prefix over processing order of the number of non-empty children
    for (int r = tid; r < m; r += nthr) {
      const int* c = &S.cc[S.byrank[r] * 4];
      S.pre[r] = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0);
    }
    if (tid == 0) { s_rstar = m - 1; s_expand = 0; }
    __syncthreads();
    {
      const int tot = qt_scan_array(S.pre, m, ws);   // exclusive in place ...
      (void)tot;
      for (int r = tid; r < m; r += nthr) {          // ... turn into inclusive
        const int* c = &S.cc[S.byrank[r] * 4];
        S.pre[r] += (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0);
      }
      __syncthreads();
    }
    if (largest) {   // :732 stop right after the split that reaches N
      for (int r = tid; r < m; r += nthr) {
        const bool now = (n + S.pre[r] - (r + 1)) >= N;
        const bool before = (r > 0) && ((n + S.pre[r - 1] - r) >= N);
        if (now && !before) s_rstar = r;
      }
      __syncthreads();
    }
    const int rstar = s_rstar;
    const int TC = (m > 0) ? S.pre[rstar] : 0;   // children created this round
    // (4) new list index of untouched nodes
    for (int i = tid; i < n; i += nthr) {
      const int r = S.rk[i];
      S.npos[i] = (r < 0 || r > rstar) ? 1 : 0;
    }
    __syncthreads();
    const int nUn = qt_scan_array(S.npos, n, ws);
    for (int i = tid; i < n; i += nthr) {
      const int r = S.rk[i];
      if (r < 0 || r > rstar) {
        const int p = TC + S.npos[i];
        S.npos[i] = p;
        nbox[p] = box[i];
        ncnt[p] = cnt[i];
        S.rk[i] = -1;
        S.cpos[i * 4 + 0] = S.cpos[i * 4 + 1] = S.cpos[i * 4 + 2] = S.cpos[i * 4 + 3] = p | QT_UNTOUCHED;
      }
    }
    // (5) children: block of rank r starts at TC - pre[r]; inside the block n4,n3,n2,n1 (non-empty only)
    int myexp = 0;
    for (int r = tid; r <= rstar && r < m; r += nthr) {
      const int i = S.byrank[r];
      const short4 b = box[i];
      const int* c = &S.cc[i * 4];
      const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);
      int p = TC - S.pre[r];
#pragma unroll
      for (int q = 3; q >= 0; --q) {
        if (c[q] > 0) {
          S.cpos[i * 4 + q] = p;
          short4 nb;
          nb.x = (q & 1) ? mx : b.x;
          nb.z = (q & 1) ? b.z : mx;
          nb.y = (q & 2) ? my : b.y;
          nb.w = (q & 2) ? b.w : my;
          nbox[p] = nb;
          ncnt[p] = c[q];
          myexp += (c[q] > 1);
          ++p;
        }
      }
    }
    if (myexp) atomicAdd(&s_expand, myexp);
    for (int i = tid; i < (TC + nUn) * 4; i += nthr) S.cc2[i] = 0;
    __syncthreads();
    // (6) re-home the keypoints; the same pass counts the quadrant populations inside the NEW nodes, i.e. the child
    //     populations the next round needs (one pass over the keypoints per round instead of two)
    // cpos[p][q] holds the new list index for quadrant q of node p -- for untouched nodes the node's own new index in
    // all four entries, flagged -- so the loop needs no branch on the node state and one box load: the child's box, and
    // with it the quadrant the keypoint will fall into NEXT round, follows from the parent's box and q.
    for (int k = tid; k < K; k += nthr) {
      const int p = qnode[k];
      const unsigned kp = qkp[k];
      const short4 b = box[p];
      const int x = kp_x(kp) - FAST_BORDER, y = kp_y(kp) - FAST_BORDER;
      const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);
      const int qx = (x < mx) ? 0 : 1, qy = (y < my) ? 0 : 1;
      const int q = qx | (qy << 1);
      const int v = S.cpos[p * 4 + q];
      const int np = v & 0x3fffffff;
      int q2 = q;
      if (!(v & QT_UNTOUCHED)) {   // the node was split: quadrant inside child q (DivideNode :478-534 of the child)
        const int cx0 = qx ? mx : b.x, cx1 = qx ? b.z : mx, cy0 = qy ? my : b.y, cy1 = qy ? b.w : my;
        const int cmx = cx0 + ((cx1 - cx0 + 1) >> 1), cmy = cy0 + ((cy1 - cy0 + 1) >> 1);
        q2 = ((x < cmx) ? 0 : 1) | ((y < cmy) ? 0 : 2);
      }
      qnode[k] = np;
      atomicAdd(&S.cc2[np * 4 + q2], 1);
    }
    { int* t = S.cc; S.cc = S.cc2; S.cc2 = t; }
    n = TC + nUn;
    const int nToExpand = s_expand;
    cur ^= 1;
    // (7) termination logic (:671-675, :736-737)
    if (n >= N || n == prevSize) finish = true;
    else if (!largest && (n + nToExpand * 3) > N) largest = true;
    __syncthreads();
  }

  // ---- best keypoint of every node (:746-762), output in list order --------------------------------------
  unsigned* best = (unsigned*)S.cc;
  for (int i = tid; i < n; i += nthr) best[i] = 0;
  __syncthreads();
  for (int k = tid; k < K; k += nthr)
    atomicMax(&best[qnode[k]], ((unsigned)kp_r(qkp[k]) << 24) | (0xffffffu - (unsigned)k));
  __syncthreads();
  unsigned* out = sel + (size_t)f * sel_per_frame + lt.sel_off[l];
  for (int i = tid; i < n; i += nthr) out[i] = qkp[0xffffffu - (best[i] & 0xffffffu)];
  if (tid == 0) selcnt[(size_t)f * nlev + l] = n;
}

#define B200_QT_ARGS LevelTab lt, const CellDesc* __restrict__ cells, const unsigned* __restrict__ cand,             \
                     const int* __restrict__ cellcnt, int ncells, int slots_per_frame, QtScratchView sc, int qt_cap,   \
                     unsigned* __restrict__ sel, int* __restrict__ selcnt, int* __restrict__ candcnt, int sel_per_frame, \
                     int level_begin, int kp_smem_cap
#define B200_QT_PASS lt, cells, cand, cellcnt, ncells, slots_per_frame, sc, qt_cap, sel, selcnt, candcnt, sel_per_frame, \
                     level_begin, kp_smem_cap
__global__ void __launch_bounds__(QT_THREADS) k_quadtree(B200_QT_ARGS) { quadtree_body(B200_QT_PASS); }
__global__ void __launch_bounds__(QT_THREADS, 3) k_quadtree_o3(B200_QT_ARGS) { quadtree_body(B200_QT_PASS); }
__global__ void __launch_bounds__(QT_THREADS, 4) k_quadtree_o4(B200_QT_ARGS) { quadtree_body(B200_QT_PASS); }
#undef B200_QT_ARGS
#undef B200_QT_PASS

// ---------------------------------------------------------------------------------------------------
// K4  GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101 (call site :1094-1095).
//     Integer kernel [18,34,48,56,48,34,18]/256 in both passes, (v + 32768) >> 16.
//     Tile 64x16 outputs per CTA; (64+6)x(16+6) source halo in shared memory; horizontal pass into u16.
// ---------------------------------------------------------------------------------------------------
constexpr int BL_TW = 64, BL_TH = 16;

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
  return p;
}

__global__ void __launch_bounds__(256) k_blur7(const uint8_t* __restrict__ src, int spitch, size_t sfs,
                                               uint8_t* __restrict__ dst, int dpitch, size_t dfs, int w, int h) {
  __shared__ uint8_t t[(BL_TH + 6)][BL_TW + 8];
  __shared__ unsigned short hb[(BL_TH + 6)][BL_TW];
  const uint8_t* s = src + (size_t)blockIdx.z * sfs;
  uint8_t* d = dst + (size_t)blockIdx.z * dfs;
  const int x0 = blockIdx.x * BL_TW, y0 = blockIdx.y * BL_TH;
  for (int i = threadIdx.x; i < (BL_TH + 6) * (BL_TW + 6); i += 256) {
    const int ty = i / (BL_TW + 6), tx = i - ty * (BL_TW + 6);
    const int sy = reflect101(y0 + ty - 3, h), sx = reflect101(x0 + tx - 3, w);
    t[ty][tx] = s[(size_t)sy * spitch + sx];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (BL_TH + 6) * BL_TW; i += 256) {
    const int ty = i / BL_TW, tx = i - ty * BL_TW;
    const uint8_t* r = &t[ty][tx];
    hb[ty][tx] = (unsigned short)(18 * (r[0] + r[6]) + 34 * (r[1] + r[5]) + 48 * (r[2] + r[4]) + 56 * r[3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BL_TH * BL_TW; i += 256) {
    const int ty = i / BL_TW, tx = i - ty * BL_TW;
    const int x = x0 + tx, y = y0 + ty;
    if (x < w && y < h) {
      const unsigned v = 18u * (hb[ty][tx] + hb[ty + 6][tx]) + 34u * (hb[ty + 1][tx] + hb[ty + 5][tx]) +
                         48u * (hb[ty + 2][tx] + hb[ty + 4][tx]) + 56u * hb[ty + 3][tx];
      d[(size_t)y * dpitch + x] = (uint8_t)((v + 32768u) >> 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K5  orientation (IC_Angle :59-88 on the raw level) + steered BRIEF (computeOrbDescriptor :92-131 on the
//     blurred level) + final KeyPoint assembly (:846-856, :1104-1110).  One warp per keypoint.
// ---------------------------------------------------------------------------------------------------
struct OrientTab {
  int umax[HALF_PATCH_SIZE + 1];
};

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {   // cv::fastAtan2, SURVEY App. A.5 (no FMA)
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float P1 = __fmul_rn(0.9997878412794807f, scale), P3 = __fmul_rn(-0.3258083974640975f, scale),
              P5 = __fmul_rn(0.1555786518463281f, scale), P7 = __fmul_rn(-0.04432655554792128f, scale);
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, 2.2204460492503131e-16f));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(P7, c2), P5), c2), P3), c2), P1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, 2.2204460492503131e-16f));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(P7, c2), P5), c2), P3), c2), P1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

constexpr int OD_WARPS = 8;

__global__ void __launch_bounds__(OD_WARPS * 32) k_orient_desc(LevelTab lt, OrientTab ot, PyrView raw, PyrView blr,
                                                               const signed char* __restrict__ pattern,
                                                               const unsigned* __restrict__ sel,
                                                               const int* __restrict__ selcnt, int sel_per_frame,
                                                               OrbxKeyPoint* __restrict__ kps,
                                                               uint8_t* __restrict__ desc, int* __restrict__ nout,
                                                               int cap) {
  __shared__ int spat[256];   // 512 points x (int8 x, int8 y) = 1024 B
  for (int i = threadIdx.x; i < 256; i += blockDim.x) spat[i] = reinterpret_cast<const int*>(pattern)[i];
  __syncthreads();
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int j = blockIdx.x * OD_WARPS + (threadIdx.x >> 5);
  const int* sc = selcnt + (size_t)f * lt.nlevels;
  int l = -1, idx = 0, acc = 0;
  for (int q = 0; q < lt.nlevels; ++q) {
    const int c = sc[q];
    if (l < 0 && j < acc + c) { l = q; idx = j - acc; }
    acc += c;
  }
  if (j == 0 && lane == 0) nout[f] = min(acc, cap);
  if (l < 0 || j >= cap) return;
  const unsigned pk = sel[(size_t)f * sel_per_frame + lt.sel_off[l] + idx];
  const int x = kp_x(pk), y = kp_y(pk);
  // ---- IC_Angle: lane <-> column u = lane-15 (31 columns), loop rows ----
  const uint8_t* rc = raw.p[l] + (size_t)f * raw.fstride[l] + (size_t)y * raw.pitch[l] + x;
  const int u = lane - HALF_PATCH_SIZE, au = abs(u);
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int rp = raw.pitch[l];
    m10 = u * rc[u];
#pragma unroll
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
      if (au <= ot.umax[v]) {
        const int vp = rc[u + v * rp], vm = rc[u - v * rp];
        m10 += u * (vp + vm);
        m01 += v * (vp - vm);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  const float angle = fast_atan2_deg((float)m01, (float)m10);
  // ---- steered BRIEF ----
  const float factorPI = (float)(3.14159265358979323846 / 180.f);
  const float ang = __fmul_rn(angle, factorPI);
  // cos()/sin() of the reference bind to glibc's cosf/sinf, which are NOT correctly rounded: run the same algorithm
  const float a = b200_cosf(ang), b = b200_sinf(ang);
  const uint8_t* bc = blr.p[l] + (size_t)f * blr.fstride[l] + (size_t)y * blr.pitch[l] + x;
  const int bp = blr.pitch[l];
  unsigned word = 0;
  unsigned* dout = reinterpret_cast<unsigned*>(desc + ((size_t)f * cap + j) * 32);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int pr = k * 32 + lane;                 // test index; bit pr of the descriptor
    const int w0 = spat[pr];                       // bytes: x0,y0,x1,y1
    const float x0 = (float)(signed char)(w0 & 0xff), y0 = (float)(signed char)((w0 >> 8) & 0xff);
    const float x1 = (float)(signed char)((w0 >> 16) & 0xff), y1 = (float)(signed char)((w0 >> 24) & 0xff);
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
    const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
    const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int t0 = bc[r0 * bp + c0], t1 = bc[r1 * bp + c1];
    const unsigned bits = __ballot_sync(0xffffffffu, t0 < t1);
    if (lane == k) word = bits;
  }
  if (lane < 8) dout[lane] = word;
  if (lane == 0) {
    OrbxKeyPoint kp;
    const float s = lt.sf[l];
    kp.x = (l != 0) ? __fmul_rn((float)x, s) : (float)x;
    kp.y = (l != 0) ? __fmul_rn((float)y, s) : (float)y;
    kp.size = lt.kp_size[l];
    kp.angle = angle;
    kp.response = (float)kp_r(pk);
    kp.octave = l;
    kp.class_id = -1;
    kps[(size_t)f * cap + j] = kp;
  }
}

// K5, second formulation (bit 0 of experimental_mask() in orbx.cu, B200ORB_EXPERIMENTAL): the same arithmetic with fewer
// instructions -- round-2 ncu capture (profiles/r02_ncu_v4_summary.txt): the first formulation is issue-bound (79 %
// issue-active, 898 warp-instructions per keypoint) with the XU pipe (I2F / F2I conversions) at 52 %.
//   * a warp handles KPW keypoints, so the per-CTA set-up (pattern into shared memory, level offsets) is paid once per
//     8 * KPW keypoints instead of once per 8;
//   * the pattern sits in shared memory as float4 (x0, y0, x1, y1) per test: one LDS.128 instead of LDS.32 + 4 I2F.S8;
//   * the level of keypoint j comes from one ballot over the per-level inclusive counts (lane q holds level q) instead
//     of an 8-iteration scan per warp;
//   * IC_Angle walks two row pointers (+pitch / -pitch) instead of rebuilding a 64-bit address from u + v * pitch per
//     load, tests the circular mask as v <= vmax(|u|) (umax is non-increasing in v -- checked at create -- so
//     {v : |u| <= umax[v]} is the prefix 1..vmax), and accumulates sum(vp + vm) with one 3-input add per row, the
//     multiplication by u once at the end (exact: integers).
// Results are identical bit for bit (integer moments; the float sequence of the steered BRIEF is unchanged).
constexpr int OD_KPW = 4;   // keypoints per warp

__global__ void __launch_bounds__(OD_WARPS * 32, 8) k_orient_desc2(LevelTab lt, OrientTab ot, PyrView raw, PyrView blr,
                                                                const float4* __restrict__ patf /* 256 tests */,
                                                                const unsigned* __restrict__ sel,
                                                                const int* __restrict__ selcnt, int sel_per_frame,
                                                                OrbxKeyPoint* __restrict__ kps,
                                                                uint8_t* __restrict__ desc, int* __restrict__ nout,
                                                                int cap) {
  __shared__ float4 spat[256];   // test pr: points 2 pr and 2 pr + 1 of the pattern, as floats
  for (int i = threadIdx.x; i < 256; i += OD_WARPS * 32) spat[i] = __ldg(patf + i);
  const int f = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // inclusive keypoint counts per level: lane q holds levels 0..q (lanes >= nlevels hold the total)
  const int cq = (lane < lt.nlevels) ? selcnt[(size_t)f * lt.nlevels + lane] : 0;
  const int incl = warp_incl_scan(cq, lane);
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (blockIdx.x == 0 && threadIdx.x == 0) nout[f] = min(total, cap);
  const int u = lane - HALF_PATCH_SIZE, au = abs(u);
  int vmaxu = 0;   // rows v = 1 .. vmaxu of column u lie inside the circular patch
#pragma unroll
  for (int v = 1; v <= HALF_PATCH_SIZE; ++v) vmaxu += (au <= ot.umax[v]) ? 1 : 0;
  __syncthreads();
  const int jmax = min(total, cap);
#pragma unroll 1
  for (int i = 0; i < OD_KPW; ++i) {
    const int j = (blockIdx.x * OD_KPW + i) * OD_WARPS + w;
    if (j >= jmax) break;   // warp-uniform
    // level of keypoint j: the first level whose inclusive count exceeds j
    const int l = __popc(__ballot_sync(0xffffffffu, lane < lt.nlevels && incl <= j));
    const int before = __shfl_sync(0xffffffffu, incl, max(l - 1, 0));
    const int idx = j - ((l > 0) ? before : 0);
    const unsigned pk = sel[(size_t)f * sel_per_frame + lt.sel_off[l] + idx];
    const int x = kp_x(pk), y = kp_y(pk);
    // ---- IC_Angle: lane <-> column u = lane-15 (31 columns), rows +-v through two walking pointers ----
    const ptrdiff_t rp = raw.pitch[l];
    int s = 0, m01 = 0;
    if (lane < 31) {
      const uint8_t* pp = raw.p[l] + (size_t)f * raw.fstride[l] + (size_t)y * rp + (x + u);
      const uint8_t* pm = pp;
      s = *pp;
#pragma unroll
      for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        pp += rp; pm -= rp;
        if (v <= vmaxu) {
          const int vp = *pp, vm = *pm;
          s += vp + vm;
          m01 += v * (vp - vm);
        }
      }
    }
    int m10 = u * s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m10 += __shfl_xor_sync(0xffffffffu, m10, o);
      m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    // ---- steered BRIEF ----
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    const float a = b200_cosf(ang), b = b200_sinf(ang);   // glibc's cosf / sinf (not correctly rounded): same algorithm
    const int bp = blr.pitch[l];
    const uint8_t* bc = blr.p[l] + (size_t)f * blr.fstride[l] + (size_t)y * bp + x;
    unsigned word = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 t = spat[k * 32 + lane];   // test k*32+lane -> bit (k*32+lane) of the descriptor
      const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(t.x, b), __fmul_rn(t.y, a)));
      const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(t.x, a), __fmul_rn(t.y, b)));
      const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(t.z, b), __fmul_rn(t.w, a)));
      const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(t.z, a), __fmul_rn(t.w, b)));
      const int t0 = bc[r0 * bp + c0], t1 = bc[r1 * bp + c1];
      const unsigned bits = __ballot_sync(0xffffffffu, t0 < t1);
      if (lane == k) word = bits;
    }
    if (lane < 8) reinterpret_cast<unsigned*>(desc + ((size_t)f * cap + j) * 32)[lane] = word;
    if (lane == 0) {
      OrbxKeyPoint kp;
      const float sc = lt.sf[l];
      kp.x = (l != 0) ? __fmul_rn((float)x, sc) : (float)x;
      kp.y = (l != 0) ? __fmul_rn((float)y, sc) : (float)y;
      kp.size = lt.kp_size[l];
      kp.angle = angle;
      kp.response = (float)kp_r(pk);
      kp.octave = l;
      kp.class_id = -1;
      kps[(size_t)f * cap + j] = kp;
    }
  }
}

// K4 (fast path): same arithmetic, register sliding window.  One thread owns 4 adjacent columns and walks down
// BLS_ROWS rows of a strip.  Per row it reads the 10 source bytes as three aligned 32-bit words, builds the byte pairs
// P_k = b[k] | b[k+2] << 16 with one funnel shift + mask each, and evaluates the horizontal pass on two pixels per
// register (every partial sum stays below 2^16, so the halves never carry into each other).  The last 7 rows of
// horizontal sums live in a statically rotated register ring for the vertical pass -- no shared memory, no
// intermediate plane.  Requires 4-byte aligned rows (true for the internal planes; level 0 falls back to k_blur7 when
// the caller's buffer is not aligned).
constexpr int BLS_ROWS = 35;   // multiple of 7: the ring rotation is unrolled by 7

struct HRow { unsigned h02, h13; };   // horizontal sums of pixels (0,2) and (1,3) of the thread's 4 columns

__device__ __forceinline__ HRow blur_hrow_words(unsigned w0, unsigned w1, unsigned w2) {
  // w0..w2 = bytes x0-4 .. x0+7.  P[k] = byte(k+1) | byte(k+3) << 16, i.e. source pixels x0-3+k and x0-1+k
  unsigned P[8];
  P[0] = __funnelshift_r(w0, w1, 8) & 0x00ff00ffu;    // bytes 1,3
  P[1] = __funnelshift_r(w0, w1, 16) & 0x00ff00ffu;   // bytes 2,4
  P[2] = __funnelshift_r(w0, w1, 24) & 0x00ff00ffu;   // bytes 3,5
  P[3] = w1 & 0x00ff00ffu;                            // bytes 4,6
  P[4] = __funnelshift_r(w1, w2, 8) & 0x00ff00ffu;    // bytes 5,7
  P[5] = __funnelshift_r(w1, w2, 16) & 0x00ff00ffu;   // bytes 6,8
  P[6] = __funnelshift_r(w1, w2, 24) & 0x00ff00ffu;   // bytes 7,9
  P[7] = w2 & 0x00ff00ffu;                            // bytes 8,10
  HRow r;   // output pixel i uses source bytes i+1 .. i+7; pair (0,2) = P[0..6], pair (1,3) = P[1..7]
  r.h02 = 18u * (P[0] + P[6]) + 34u * (P[1] + P[5]) + 48u * (P[2] + P[4]) + 56u * P[3];
  r.h13 = 18u * (P[1] + P[7]) + 34u * (P[2] + P[6]) + 48u * (P[3] + P[5]) + 56u * P[4];
  return r;
}

// The same horizontal sums with byte dot products: output pixel i needs the 7 bytes i+1 .. i+7 of the 12-byte window;
// bytes i+1..i+4 and i+5..i+8 are two funnel shifts of the three words (none for i = 3), and two IDP.4A against the
// coefficient words (18,34,48,56) / (48,34,18,0) give the exact integer sum.  h[0..3] are at most 255*256 = 65280.
__device__ __forceinline__ void blur_hrow_dp(unsigned w0, unsigned w1, unsigned w2, unsigned h[4]) {
  constexpr unsigned cA = 18u | (34u << 8) | (48u << 16) | (56u << 24), cB = 48u | (34u << 8) | (18u << 16);
  h[0] = __dp4a(__funnelshift_r(w0, w1, 8), cA, __dp4a(__funnelshift_r(w1, w2, 8), cB, 0u));
  h[1] = __dp4a(__funnelshift_r(w0, w1, 16), cA, __dp4a(__funnelshift_r(w1, w2, 16), cB, 0u));
  h[2] = __dp4a(__funnelshift_r(w0, w1, 24), cA, __dp4a(__funnelshift_r(w1, w2, 24), cB, 0u));
  h[3] = __dp4a(w1, cA, __dp4a(w2, cB, 0u));
}

// interior column groups only (x0 >= 4 and x0 + 8 <= w): three aligned word loads, no border logic
__device__ __forceinline__ HRow blur_hrow(const uint8_t* __restrict__ row, int x0) {
  const unsigned* p = reinterpret_cast<const unsigned*>(row + x0 - 4);
  return blur_hrow_words(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

// border column groups: the same three words assembled through BORDER_REFLECT_101
__device__ __forceinline__ HRow blur_hrow_border(const uint8_t* __restrict__ row, int x0, int w) {
  unsigned b[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) b[i] = __ldg(row + reflect101(x0 - 4 + i, w));
  return blur_hrow_words(b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24), b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24),
                         b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24));
}

struct BlurTile { short level, x0, y0, pad; };   // one warp-tile: 128 columns x BLS_ROWS rows of one level

// All levels in ONE launch (a thread walks 35+6 rows sequentially, so a per-level launch is bounded below by that
// latency chain; one launch lets the small levels hide inside the big ones).  One warp per tile, blockIdx.y = frame.
__global__ void __launch_bounds__(32) k_blur7_strip(PyrView src, PyrView dstv, const BlurTile* __restrict__ tiles) {
  const BlurTile t = tiles[blockIdx.x];
  const int l = t.level, w = src.w[l], h = src.h[l];
  const int x0 = t.x0 + threadIdx.x * 4;
  if (x0 < 4 || x0 + 8 > w) return;   // border column groups (<= 3 per row) are done by k_blur7_edges
  const int spitch = src.pitch[l], dpitch = dstv.pitch[l];
  const uint8_t* s = src.p[l] + (size_t)blockIdx.y * src.fstride[l];
  uint8_t* d = dstv.p[l] + (size_t)blockIdx.y * dstv.fstride[l];
  const int y0 = t.y0, y1 = min(h, y0 + BLS_ROWS);
  unsigned ring[7][4];   // unpacked horizontal sums (pixels 0..3) of 7 consecutive rows, slot = row phase
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const HRow r = blur_hrow(s + (size_t)reflect101(y0 - 3 + j, h) * spitch, x0);
    ring[j][0] = r.h02 & 0xffffu; ring[j][2] = r.h02 >> 16; ring[j][1] = r.h13 & 0xffffu; ring[j][3] = r.h13 >> 16;
  }
  for (int yb = y0; yb < y1; yb += 7) {
    if (yb + 7 <= y1) {
      // full block: all 21 source words of the 7 new rows are requested before the first one is used, so a warp keeps
      // 7 rows (2.7 KB) in flight instead of one -- the walk is latency-bound otherwise
      unsigned W[7][3];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const unsigned* p = reinterpret_cast<const unsigned*>(s + (size_t)reflect101(yb + k + 3, h) * spitch + x0 - 4);
        W[k][0] = __ldg(p); W[k][1] = __ldg(p + 1); W[k][2] = __ldg(p + 2);
      }
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        blur_hrow_dp(W[k][0], W[k][1], W[k][2], ring[(k + 6) % 7]);
        unsigned out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned v = 18u * (ring[k % 7][i] + ring[(k + 6) % 7][i]) + 34u * (ring[(k + 1) % 7][i] + ring[(k + 5) % 7][i]) +
                             48u * (ring[(k + 2) % 7][i] + ring[(k + 4) % 7][i]) + 56u * ring[(k + 3) % 7][i];
          out |= ((v + 32768u) >> 16) << (8 * i);
        }
        *reinterpret_cast<unsigned*>(d + (size_t)(yb + k) * dpitch + x0) = out;   // dst pitch is a multiple of 16
      }
      continue;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int y = yb + k;
      if (y < y1) {
        // rows y-3 .. y+2 sit in slots (k+0)%7 .. (k+5)%7; the new row y+3 goes to slot (k+6)%7
        const HRow r = blur_hrow(s + (size_t)reflect101(y + 3, h) * spitch, x0);
        ring[(k + 6) % 7][0] = r.h02 & 0xffffu; ring[(k + 6) % 7][2] = r.h02 >> 16;
        ring[(k + 6) % 7][1] = r.h13 & 0xffffu; ring[(k + 6) % 7][3] = r.h13 >> 16;
        unsigned out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned v = 18u * (ring[k % 7][i] + ring[(k + 6) % 7][i]) + 34u * (ring[(k + 1) % 7][i] + ring[(k + 5) % 7][i]) +
                             48u * (ring[(k + 2) % 7][i] + ring[(k + 4) % 7][i]) + 56u * ring[(k + 3) % 7][i];
          out |= ((v + 32768u) >> 16) << (8 * i);
        }
        *reinterpret_cast<unsigned*>(d + (size_t)y * dpitch + x0) = out;   // dst pitch is a multiple of 16
      }
    }
  }
}

// Border column groups of every level (the group at x0 = 0 and the last one or two groups): same register-ring walk
// as k_blur7_strip with the horizontal pass fetched through BORDER_REFLECT_101; one thread per (group, 35-row strip).
// ~3 % of the pixels.
struct BlurEdge { short level, x0, y0, pad; };

__global__ void __launch_bounds__(64) k_blur7_edges(PyrView src, PyrView dstv, const BlurEdge* __restrict__ edges, int nedges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nedges) return;
  const BlurEdge e = edges[i];
  const int l = e.level, x0 = e.x0, w = src.w[l], h = src.h[l], spitch = src.pitch[l], dpitch = dstv.pitch[l];
  const uint8_t* s = src.p[l] + (size_t)blockIdx.y * src.fstride[l];
  uint8_t* d = dstv.p[l] + (size_t)blockIdx.y * dstv.fstride[l];
  const int y0 = e.y0, y1 = min(h, y0 + BLS_ROWS);
  unsigned ring[7][4];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const HRow r = blur_hrow_border(s + (size_t)reflect101(y0 - 3 + j, h) * spitch, x0, w);
    ring[j][0] = r.h02 & 0xffffu; ring[j][2] = r.h02 >> 16; ring[j][1] = r.h13 & 0xffffu; ring[j][3] = r.h13 >> 16;
  }
  for (int yb = y0; yb < y1; yb += 7) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int y = yb + k;
      if (y < y1) {
        const HRow r = blur_hrow_border(s + (size_t)reflect101(y + 3, h) * spitch, x0, w);
        ring[(k + 6) % 7][0] = r.h02 & 0xffffu; ring[(k + 6) % 7][2] = r.h02 >> 16;
        ring[(k + 6) % 7][1] = r.h13 & 0xffffu; ring[(k + 6) % 7][3] = r.h13 >> 16;
        unsigned out = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned v = 18u * (ring[k % 7][q] + ring[(k + 6) % 7][q]) + 34u * (ring[(k + 1) % 7][q] + ring[(k + 5) % 7][q]) +
                             48u * (ring[(k + 2) % 7][q] + ring[(k + 4) % 7][q]) + 56u * ring[(k + 3) % 7][q];
          out |= ((v + 32768u) >> 16) << (8 * q);
        }
        *reinterpret_cast<unsigned*>(d + (size_t)y * dpitch + x0) = out;
      }
    }
  }
}

// bordered level read-back for orbx_get_level (copyMakeBorder BORDER_REFLECT_101, :1136-1142)
__global__ void k_border_copy(const uint8_t* __restrict__ src, int spitch, int w, int h, uint8_t* __restrict__ dst,
                              int dpitch, int B) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w + 2 * B || y >= h + 2 * B) return;
  dst[(size_t)y * dpitch + x] = src[(size_t)reflect101(y - B, h) * spitch + reflect101(x - B, w)];
}

}  // namespace b200
