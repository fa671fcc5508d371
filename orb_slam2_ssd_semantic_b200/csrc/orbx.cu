// orbx.cu -- host side + C-ABI of the B200 ORB extractor (reference: src/ORBextractor.cc,
// include/ORBextractor.h).  Kernels live in orbx_kernels.cuh.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <new>
#include <vector>

#include "orbx_kernels.cuh"
#include "orbx_host.h"

namespace b200 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libb200orb has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return B200ORB_ENOGPU;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (0..%d)", device, n - 1);
    return B200ORB_EINVAL;
  }
  return B200ORB_OK;
}

static const signed char kPatternHost[1024] = {
#include "../../include/orb_pattern_31.inc"
};

// Kernel selection.  `exp_mask` switches the formulations added in round 2 (bit 0 = k_orient_desc2, bit 1 = second tile
// staging of k_fast_cells; the first ones stay available for A/B); `fast_wpc` (FAST cells
// = warps per CTA: 1, 2, 4, 8) and `qt_minb` (quad-tree register budget for 2, 3 or 4 CTAs per SM) only change launch
// shapes.  Defaults below; the environment (B200ORB_EXPERIMENTAL, B200ORB_FAST_WPC, B200ORB_QT_MINB) overrides them at
// load, b200orb_set_tuning() at run time (tools/tune_extractor.py); b200orb_experimental() / b200orb_get_tuning() report.
// Defaults = what the B200 runs of round 2 ended on (profiles/r02_notes.md): both formulations passed the whole GPU parity
// suite (B200ORB_EXPERIMENTAL=3: 77 passed), tools/tune_extractor.py found every setting byte-identical to (0, 8, 2), the
// quad-tree 7 - 19 % faster with 4 CTAs per SM and the FAST CTA size irrelevant (within 1 %: the suite's 8 stays).
// B200ORB_EXPERIMENTAL=0 B200ORB_QT_MINB=2 restores round 1's kernels.
constexpr int kExperimentalDefault = 3, kFastWpcDefault = 8, kQtMinbDefault = 4;
constexpr int EXP_ORIENT2 = 1, EXP_FAST_STAGE2 = 2;
struct Tuning { int exp_mask, fast_wpc, qt_minb; };
static inline bool valid_wpc(int v) { return v == 1 || v == 2 || v == 4 || v == 8; }
static inline bool valid_minb(int v) { return v >= 2 && v <= 4; }
static Tuning& tuning() {
  static Tuning t = [] {
    Tuning d{kExperimentalDefault, kFastWpcDefault, kQtMinbDefault};
    if (const char* e = getenv("B200ORB_EXPERIMENTAL")) d.exp_mask = atoi(e);
    if (const char* e = getenv("B200ORB_FAST_WPC")) { const int v = atoi(e); if (valid_wpc(v)) d.fast_wpc = v; }
    if (const char* e = getenv("B200ORB_QT_MINB")) { const int v = atoi(e); if (valid_minb(v)) d.qt_minb = v; }
    return d;
  }();
  return t;
}
static inline int experimental_mask() { return tuning().exp_mask; }

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

}  // namespace b200

using namespace b200;

// =====================================================================================================
// handle
// =====================================================================================================
orbx::orbx() {}

int orbx::init(const OrbxParams& p, int dev) {
  prm = p;
  device = dev;
  const int nl = p.nlevels;
  // ---- ctor arithmetic, src/ORBextractor.cc:404-465 (scaleFactor member is double, ctor arg float) ----
  const double scaleFactor = (double)p.scale_factor;
  sf[0] = 1.0f; sigma2[0] = 1.0f;
  for (int i = 1; i < nl; ++i) {
    sf[i] = (float)(sf[i - 1] * scaleFactor);
    sigma2[i] = sf[i] * sf[i];
  }
  for (int i = 0; i < nl; ++i) { invsf[i] = 1.0f / sf[i]; invsigma2[i] = 1.0f / sigma2[i]; }
  const float factor = (float)(1.0f / scaleFactor);
  float nDesired = p.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; ++l) {
    nfeat[l] = cv_round_f(nDesired);
    sum += nfeat[l];
    nDesired *= factor;
  }
  nfeat[nl - 1] = std::max(p.nfeatures - sum, 0);
  int v, v0;
  const int vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
  const int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
  const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
  for (v = 0; v <= vmax; ++v) otab.umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
  for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
    while (otab.umax[v0] == otab.umax[v0 + 1]) ++v0;
    otab.umax[v] = v0;
    ++v0;
  }
  B200_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  for (v = 1; v <= HALF_PATCH_SIZE; ++v)   // k_orient_desc2 tests the circular mask as v <= vmax(|u|)
    if (otab.umax[v] > otab.umax[v - 1]) { set_error("umax table is not non-increasing"); return B200ORB_EINVAL; }
  B200_CUDA(cudaMalloc(&d_pattern, 1024));
  B200_CUDA(cudaMemcpyAsync(d_pattern, kPatternHost, 1024, cudaMemcpyHostToDevice, stream));
  float patf[1024];
  for (int i = 0; i < 1024; ++i) patf[i] = (float)kPatternHost[i];
  B200_CUDA(cudaMalloc(&d_patf, sizeof(patf)));
  B200_CUDA(cudaMemcpyAsync(d_patf, patf, sizeof(patf), cudaMemcpyHostToDevice, stream));
  B200_CUDA(cudaStreamSynchronize(stream));
  return B200ORB_OK;
}

void orbx::free_geometry() {
  auto F = [](void* p) { if (p) cudaFree(p); };
  F(d_raw); F(d_blur); F(d_cells); F(d_cand); F(d_cellcnt); F(d_qkp); F(d_qnode); F(d_sel); F(d_selcnt);
  F(d_candcnt); F(d_kps); F(d_desc); F(d_n); F(d_xt); F(d_yt); F(d_xg); F(d_tmp); F(d_blur_tiles); F(d_blur_edges);
  d_blur_tiles = nullptr; d_blur_edges = nullptr;
  d_raw = d_blur = nullptr; d_cells = nullptr; d_cand = nullptr; d_cellcnt = nullptr; d_qkp = nullptr;
  d_qnode = nullptr; d_sel = nullptr; d_selcnt = nullptr; d_candcnt = nullptr; d_kps = nullptr; d_desc = nullptr;
  d_n = nullptr; d_xt = d_yt = nullptr; d_xg = nullptr; d_tmp = nullptr; tmp_bytes = 0;
  if (h_stage) cudaFreeHost(h_stage);
  h_stage = nullptr; stage_bytes = 0;
  rows = cols = maxF = 0;
}

orbx::~orbx() {
  DeviceGuard g(device);
  free_geometry();
  if (d_pattern) cudaFree(d_pattern);
  if (d_patf) cudaFree(d_patf);
  if (stream) cudaStreamDestroy(stream);
}

// (re)build every geometry-dependent table and buffer for F frames of rows x cols
int orbx::ensure_geometry(int r, int c, int F) {
  if (r == rows && c == cols && F <= maxF) return B200ORB_OK;
  if (r > 4095 || c > 4095) { set_error("image %dx%d exceeds the 4095-px packing limit", c, r); return B200ORB_EINVAL; }
  const int keepF = (r == rows && c == cols) ? std::max(F, maxF) : F;
  B200_CUDA(cudaStreamSynchronize(stream));
  free_geometry();
  const int nl = prm.nlevels;
  // level sizes (:1121-1122)
  size_t off = 0;
  for (int l = 0; l < nl; ++l) {
    lw[l] = cv_round_f((float)c * invsf[l]);
    lh[l] = cv_round_f((float)r * invsf[l]);
    if (lw[l] <= 2 * EDGE_THRESHOLD - 6 + 0 || lh[l] <= 2 * EDGE_THRESHOLD - 6 + 0) {
      set_error("level %d is %dx%d: too small for the 16-px FAST border (reference would crash)", l, lw[l], lh[l]);
      return B200ORB_EGEOM;
    }
    lpitch[l] = align_up(lw[l], 16);
    loff[l] = off;
    off += (size_t)lpitch[l] * lh[l];
  }
  frame_bytes = align_up_sz(off, 256);
  // resize tables
  std::vector<int2> xt, yt;
  std::vector<int4> xg;
  resize_group_ok = true;
  for (int l = 1; l < nl; ++l) {
    xt_off[l] = (int)xt.size();
    yt_off[l] = (int)yt.size();
    auto fill = [](int dn, int sn, std::vector<int2>& t) {
      const double scale = 1.0 / ((double)dn / sn);
      for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int i = (int)std::floor(f);
        f -= (float)i;
        if (i < 0) { i = 0; f = 0.f; }
        if (i >= sn - 1) { i = sn - 1; f = 0.f; }
        const int a0 = (short)cv_round_f((1.f - f) * 2048.f), a1 = (short)cv_round_f(f * 2048.f);
        t.push_back(make_int2(i, (a0 & 0xffff) | (a1 << 16)));
      }
    };
    fill(lw[l], lw[l - 1], xt);
    fill(lh[l], lh[l - 1], yt);
    // group table of k_resize_g
    xg_off[l] = (int)xg.size();
    for (int dx0 = 0; dx0 < lw[l]; dx0 += 4) {
      const int2* t = xt.data() + xt_off[l];
      int sx[4], cf[4];
      for (int i = 0; i < 4; ++i) { const int2 e = t[std::min(dx0 + i, lw[l] - 1)]; sx[i] = e.x; cf[i] = e.y; }
      for (int i = 1; i < 4; ++i) if (sx[i] < sx[0]) sx[i] = sx[0];   // (replicated tail entries never precede sx[0])
      int sels = 0;
      for (int i = 0; i < 4; ++i) {
        const int dlt = sx[i] - sx[0];
        if (dlt + 1 > 7) resize_group_ok = false;   // scale > 2: the byte-window path does not apply
        sels |= (((dlt & 7) | (((dlt + 1) & 7) << 4)) & 0xff) << (8 * i);
      }
      xg.push_back(make_int4(sx[0] >> 2, 8 * (sx[0] & 3), sels, 0));
      xg.push_back(make_int4(cf[0], cf[1], cf[2], cf[3]));
    }
  }
  // FAST cells (:775-815) and quad-tree constants (:545-547)
  std::vector<CellDesc> cells;
  int slots = 0, selcap = 0, qcap = 8;
  memset(&ltab, 0, sizeof(ltab));
  ltab.nlevels = nl;
  for (int l = 0; l < nl; ++l) {
    ltab.cell_begin[l] = (int)cells.size();
    ltab.slot_begin[l] = slots;
    const int minBX = FAST_BORDER, minBY = FAST_BORDER;
    const int maxBX = lw[l] - EDGE_THRESHOLD + 3, maxBY = lh[l] - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const float W = 30;
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) {
      set_error("level %d (%dx%d) has no 30-px FAST cell: the reference divides by zero here", l, lw[l], lh[l]);
      return B200ORB_EGEOM;
    }
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    for (int i = 0; i < nRows; ++i) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; ++j) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        CellDesc cd;
        cd.level = (short)l;
        cd.x0 = (short)iniX; cd.y0 = (short)iniY;
        cd.rw = (short)((int)maxX - (int)iniX); cd.rh = (short)((int)maxY - (int)iniY);
        cd.pad = 0;
        cd.slot_off = slots;
        if (cd.rw > FAST_MAX_ROI - 2 || cd.rh > FAST_MAX_ROI) { set_error("FAST cell larger than %d px", FAST_MAX_ROI); return B200ORB_EGEOM; }
        const int iw = cd.rw - 6, ih = cd.rh - 6;
        if (iw > 0 && ih > 0) slots += ((iw + 1) / 2) * ((ih + 1) / 2);
        cells.push_back(cd);
      }
    }
    const int nIni = (int)std::round(static_cast<float>(maxBX - minBX) / (maxBY - minBY));
    if (nIni <= 0) { set_error("aspect ratio of level %d makes nIni = 0 (reference divides by zero)", l); return B200ORB_EGEOM; }
    ltab.n_ini[l] = nIni;
    ltab.hx[l] = static_cast<float>(maxBX - minBX) / nIni;
    ltab.box_h[l] = maxBY - minBY;
    ltab.nfeat[l] = nfeat[l];
    ltab.sf[l] = sf[l];
    ltab.kp_size[l] = (float)(int)(PATCH_SIZE * sf[l]);
    const int capl = std::max(nfeat[l] + 3, 4 * nIni);
    ltab.sel_off[l] = selcap;
    selcap += capl;
    qcap = std::max(qcap, capl + 1);
  }
  ltab.cell_begin[nl] = (int)cells.size();
  ltab.slot_begin[nl] = slots;
  ltab.sel_off[nl] = selcap;
  ncells = (int)cells.size();
  {
    int rwm = 0, rhm = 0, det = 0;
    for (const CellDesc& cd : cells) {
      rwm = std::max(rwm, (int)cd.rw); rhm = std::max(rhm, (int)cd.rh);
      det = std::max(det, ((int)cd.rw - 6) * ((int)cd.rh - 6));
    }
    fast_clist_cap = align_up(std::max(det, 8), 8);   // every detection pixel of a cell may be a corner
    fast_tp = (rwm + 3 <= FAST_TP_SMALL) ? FAST_TP_SMALL : FAST_TP_BIG;
    fast_rows_max = rhm;
    fast_smem = (size_t)FAST_WARPS * (2 * rhm * fast_tp + 2 * FAST_QLEN + 2 * fast_clist_cap);
    if (fast_smem > 48 * 1024) {
      B200_CUDA(cudaFuncSetAttribute(k_fast_cells<FAST_TP_BIG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
      B200_CUDA(cudaFuncSetAttribute(k_fast_cells<FAST_TP_SMALL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
      B200_CUDA(cudaFuncSetAttribute(k_fast_cells<FAST_TP_BIG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
      B200_CUDA(cudaFuncSetAttribute(k_fast_cells<FAST_TP_SMALL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
    }
  }
  slots_per_frame = align_up(slots, 4);
  sel_per_frame = selcap;
  cap = selcap;
  qt_cap = align_up(qcap, 8);
  qt_smem = qt_smem_bytes(qt_cap);
  if (qt_smem > 100 * 1024) { set_error("nfeatures too large for the quad-tree kernel's shared memory"); return B200ORB_EINVAL; }
  {
    // Levels are launched in three groups with decreasing shared-memory budgets (level 0-1: 200 KB, 2-4: 112 KB,
    // 5+: 56 KB) so that small levels keep several CTAs per SM; a (frame, level) whose candidates exceed its group's
    // budget uses the global scratch instead (same code path, generic pointers).
    // Measured on B200 (profiles/r01_notes.md): staging the candidates in shared memory cuts a CTA's latency but
    // costs more in lost CTA-level concurrency (1 CTA/SM for the big levels, split launches for the small ones) than
    // it gains: 0.51 ms -> 0.74-0.83 ms per 256 frames.  All levels therefore run in ONE launch on the L2-resident
    // scratch (budget 0); the staging path stays available behind the budgets.
    const int bounds[4] = {0, nl, nl, nl};
    const size_t budget[3] = {0, 0, 0};
    size_t mx = 0;
    for (int g = 0; g < 3; ++g) {
      qt_group_lb[g] = bounds[g]; qt_group_le[g] = bounds[g + 1];
      int worst = 0;
      for (int l = bounds[g]; l < bounds[g + 1]; ++l) worst = std::max(worst, ltab.slot_begin[l + 1] - ltab.slot_begin[l]);
      size_t room = budget[g] > qt_smem ? budget[g] - qt_smem : 0;
      int kcap = (int)std::min<size_t>((size_t)worst, room / 8) & ~3;
      qt_group_kcap[g] = std::max(kcap, 0);
      qt_group_smem[g] = qt_smem + (size_t)qt_group_kcap[g] * 8;
      mx = std::max(mx, qt_group_smem[g]);
    }
    B200_CUDA(cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
    B200_CUDA(cudaFuncSetAttribute(k_quadtree_o3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
    B200_CUDA(cudaFuncSetAttribute(k_quadtree_o4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
  }

  rows = r; cols = c; maxF = keepF;
  const size_t Fz = (size_t)maxF;
  B200_CUDA(cudaMalloc(&d_raw, frame_bytes * Fz));
  B200_CUDA(cudaMalloc(&d_blur, frame_bytes * Fz));
  B200_CUDA(cudaMalloc(&d_cells, sizeof(CellDesc) * ncells));
  B200_CUDA(cudaMalloc(&d_cand, sizeof(unsigned) * slots_per_frame * Fz));
  B200_CUDA(cudaMalloc(&d_cellcnt, sizeof(int) * ncells * Fz));
  B200_CUDA(cudaMalloc(&d_qkp, sizeof(unsigned) * slots_per_frame * Fz));
  B200_CUDA(cudaMalloc(&d_qnode, sizeof(int) * slots_per_frame * Fz));
  B200_CUDA(cudaMalloc(&d_sel, sizeof(unsigned) * sel_per_frame * Fz));
  B200_CUDA(cudaMalloc(&d_selcnt, sizeof(int) * nl * Fz));
  B200_CUDA(cudaMalloc(&d_candcnt, sizeof(int) * nl * Fz));
  B200_CUDA(cudaMalloc(&d_kps, sizeof(OrbxKeyPoint) * cap * Fz));
  B200_CUDA(cudaMalloc(&d_desc, (size_t)32 * cap * Fz));
  B200_CUDA(cudaMalloc(&d_n, sizeof(int) * Fz));
  B200_CUDA(cudaMalloc(&d_xt, sizeof(int2) * std::max<size_t>(xt.size(), 1)));
  B200_CUDA(cudaMalloc(&d_yt, sizeof(int2) * std::max<size_t>(yt.size(), 1)));
  B200_CUDA(cudaMalloc(&d_xg, sizeof(int4) * std::max<size_t>(xg.size(), 1)));
  {
    std::vector<BlurTile> bt;
    for (int l = 0; l < nl; ++l)
      for (int y = 0; y < lh[l]; y += BLS_ROWS)
        for (int x = 0; x < lw[l]; x += 128) bt.push_back(BlurTile{(short)l, (short)x, (short)y, 0});
    n_blur_tiles = (int)bt.size();
    std::vector<BlurEdge> be;
    for (int l = 0; l < nl; ++l)
      for (int x = 0; x < lw[l]; x += 4)
        if (x < 4 || x + 8 > lw[l])
          for (int y = 0; y < lh[l]; y += BLS_ROWS) be.push_back(BlurEdge{(short)l, (short)x, (short)y, 0});
    n_blur_edges = (int)be.size();
    B200_CUDA(cudaMalloc(&d_blur_edges, sizeof(BlurEdge) * be.size()));
    B200_CUDA(cudaMemcpyAsync(d_blur_edges, be.data(), sizeof(BlurEdge) * be.size(), cudaMemcpyHostToDevice, stream));
    B200_CUDA(cudaMalloc(&d_blur_tiles, sizeof(BlurTile) * bt.size()));
    B200_CUDA(cudaMemcpyAsync(d_blur_tiles, bt.data(), sizeof(BlurTile) * bt.size(), cudaMemcpyHostToDevice, stream));
    B200_CUDA(cudaStreamSynchronize(stream));   // bt is a local
  }
  B200_CUDA(cudaMemcpyAsync(d_cells, cells.data(), sizeof(CellDesc) * ncells, cudaMemcpyHostToDevice, stream));
  if (!xt.empty()) B200_CUDA(cudaMemcpyAsync(d_xt, xt.data(), sizeof(int2) * xt.size(), cudaMemcpyHostToDevice, stream));
  if (!yt.empty()) B200_CUDA(cudaMemcpyAsync(d_yt, yt.data(), sizeof(int2) * yt.size(), cudaMemcpyHostToDevice, stream));
  if (!xg.empty()) B200_CUDA(cudaMemcpyAsync(d_xg, xg.data(), sizeof(int4) * xg.size(), cudaMemcpyHostToDevice, stream));
  B200_CUDA(cudaStreamSynchronize(stream));
  // blurred pyramid view never changes; raw view's level 0 may alias the caller's buffer per call
  for (int l = 0; l < nl; ++l) {
    blurv.p[l] = d_blur + loff[l]; blurv.fstride[l] = frame_bytes; blurv.pitch[l] = lpitch[l];
    blurv.w[l] = lw[l]; blurv.h[l] = lh[l];
    rawv.p[l] = d_raw + loff[l]; rawv.fstride[l] = frame_bytes; rawv.pitch[l] = lpitch[l];
    rawv.w[l] = lw[l]; rawv.h[l] = lh[l];
  }
  have_results = false;
  return B200ORB_OK;
}

// Enqueue the whole extractor for F frames whose level 0 is (d_l0, pitch0, fstride0) on `stream`.
// Enqueue the whole extractor for frames [f0, f0+F) of the batch; d_l0 = level 0 of frame f0.  Ranges of one batch may be
// enqueued one after another (the host path uploads and extracts in chunks so that PCIe and the SMs overlap).
int orbx::run(const uint8_t* d_l0, int pitch0, size_t fstride0, int F, int f0) {
  const int nl = prm.nlevels;
  // frame-range views: every per-frame array is addressed as base + frame * stride inside the kernels
  PyrView rawv = this->rawv, blurv = this->blurv;
  for (int l = 1; l < nl; ++l) rawv.p[l] += (size_t)f0 * rawv.fstride[l];
  for (int l = 0; l < nl; ++l) blurv.p[l] += (size_t)f0 * blurv.fstride[l];
  rawv.p[0] = const_cast<uint8_t*>(d_l0);
  rawv.pitch[0] = pitch0;
  rawv.fstride[0] = fstride0;
  if (f0 == 0) {   // what orbx_get_level reads back
    this->rawv.p[0] = const_cast<uint8_t*>(d_l0); this->rawv.pitch[0] = pitch0; this->rawv.fstride[0] = fstride0;
  }
  unsigned* d_cand = this->d_cand + (size_t)f0 * slots_per_frame;
  int* d_cellcnt = this->d_cellcnt + (size_t)f0 * ncells;
  unsigned* d_qkp = this->d_qkp + (size_t)f0 * slots_per_frame;
  int* d_qnode = this->d_qnode + (size_t)f0 * slots_per_frame;
  unsigned* d_sel = this->d_sel + (size_t)f0 * sel_per_frame;
  int* d_selcnt = this->d_selcnt + (size_t)f0 * nl;
  int* d_candcnt = this->d_candcnt + (size_t)f0 * nl;
  OrbxKeyPoint* d_kps = this->d_kps + (size_t)f0 * cap;
  uint8_t* d_desc = this->d_desc + (size_t)f0 * cap * 32;
  int* d_n = this->d_n + f0;
  B200_CHECK(prof_begin(F));
  // K1 pyramid
  for (int l = 1; l < nl; ++l) {
    dim3 blk(32, 8), grd((lw[l] + 127) / 128, (lh[l] + 7) / 8, F);
    // k_resize_g needs aligned source rows and 4 outputs within an 8-byte source window (scale factor <= 2)
    const bool src_aligned = prm.scale_factor <= 2.0f && (((uintptr_t)rawv.p[l - 1]) & 3) == 0 && (rawv.pitch[l - 1] & 3) == 0 && (rawv.fstride[l - 1] & 3) == 0;
    if (src_aligned && resize_group_ok)
      k_resize_g<<<grd, blk, 0, stream>>>(rawv.p[l - 1], rawv.pitch[l - 1], rawv.fstride[l - 1], lh[l - 1], rawv.p[l],
                                         rawv.pitch[l], rawv.fstride[l], lw[l], lh[l], d_xg + xg_off[l], d_yt + yt_off[l]);
    else
      k_resize<<<grd, blk, 0, stream>>>(rawv.p[l - 1], rawv.pitch[l - 1], rawv.fstride[l - 1], lw[l - 1], lh[l - 1],
                                       rawv.p[l], rawv.pitch[l], rawv.fstride[l], lw[l], lh[l], d_xt + xt_off[l],
                                       d_yt + yt_off[l]);
    ++launches;
  }
  B200_CHECK(prof_mark(ST_RESIZE + 1));
  // K2 FAST cells
  int fast_aligned = 1;   // every level's rows 4-byte aligned? (internal planes always are; level 0 may alias a caller buffer)
  for (int l = 0; l < nl; ++l)
    if ((((uintptr_t)rawv.p[l]) & 3) || (rawv.pitch[l] & 3) || (rawv.fstride[l] & 3)) fast_aligned = 0;
  if (fast_aligned && (experimental_mask() & EXP_FAST_STAGE2)) fast_aligned |= 2;   // second tile staging of k_fast_cells
  {
    // cells (= warps) per CTA: smaller CTAs (same kernel: a warp never synchronises with its neighbours) shorten the
    // tail a CTA spends waiting for its slowest cell
    const int wpc = tuning().fast_wpc;
    const dim3 grd((ncells + wpc - 1) / wpc, F);
    const size_t fast_smem_l = fast_smem / FAST_WARPS * wpc;
    // B200ORB_FAST_SWEEP=1 selects the one-pixel-per-lane sweep (development A/B switch; default: the word sweep)
    static const bool sweep4 = [] { const char* e = getenv("B200ORB_FAST_SWEEP"); return !(e && e[0] == '1'); }();
#define B200_FAST_LAUNCH(TPV, S4)                                                                                          \
  k_fast_cells<TPV, S4><<<grd, wpc * 32, fast_smem_l, stream>>>(rawv, d_cells, ncells, slots_per_frame, prm.ini_th_fast,   \
                                                                  prm.min_th_fast, d_cand, d_cellcnt, fast_aligned,        \
                                                                  fast_rows_max, fast_clist_cap)
    if (fast_tp == FAST_TP_SMALL) { if (sweep4) B200_FAST_LAUNCH(FAST_TP_SMALL, true); else B200_FAST_LAUNCH(FAST_TP_SMALL, false); }
    else { if (sweep4) B200_FAST_LAUNCH(FAST_TP_BIG, true); else B200_FAST_LAUNCH(FAST_TP_BIG, false); }
#undef B200_FAST_LAUNCH
  }
  ++launches;
  B200_CHECK(prof_mark(ST_FAST + 1));
  // K3 quad-tree: levels are launched in two groups so that each group's per-candidate arrays fit shared memory
  QtScratchView qs{d_qkp, d_qnode};
  for (int g = 0; g < 3; ++g) {
    const int lb = qt_group_lb[g], le = qt_group_le[g];
    if (le <= lb) continue;
    auto* qt_kernel = tuning().qt_minb == 4 ? k_quadtree_o4 : tuning().qt_minb == 3 ? k_quadtree_o3 : k_quadtree;
    qt_kernel<<<dim3(le - lb, F), QT_THREADS, qt_group_smem[g], stream>>>(ltab, d_cells, d_cand, d_cellcnt, ncells,
                                                                          slots_per_frame, qs, qt_cap, d_sel, d_selcnt,
                                                                          d_candcnt, sel_per_frame, lb, qt_group_kcap[g]);
    ++launches;
  }
  B200_CHECK(prof_mark(ST_QUADTREE + 1));
  if (after_select) B200_CHECK(after_select(after_select_ctx, f0, F));
  // K4 blur: one launch over every (level, 128x35 tile, frame) when all planes are 4-byte aligned
  {
    bool aligned = true;
    for (int l = 0; l < nl; ++l)
      if ((((uintptr_t)rawv.p[l]) & 3) || (rawv.pitch[l] & 3) || (rawv.fstride[l] & 3)) aligned = false;
    if (aligned) {
      k_blur7_strip<<<dim3(n_blur_tiles, F), 32, 0, stream>>>(rawv, blurv, d_blur_tiles);
      k_blur7_edges<<<dim3((n_blur_edges + 63) / 64, F), 64, 0, stream>>>(rawv, blurv, d_blur_edges, n_blur_edges);
      launches += 2;
    } else {
      for (int l = 0; l < nl; ++l) {
        dim3 grd((lw[l] + BL_TW - 1) / BL_TW, (lh[l] + BL_TH - 1) / BL_TH, F);
        k_blur7<<<grd, 256, 0, stream>>>(rawv.p[l], rawv.pitch[l], rawv.fstride[l], blurv.p[l], blurv.pitch[l],
                                         blurv.fstride[l], lw[l], lh[l]);
        ++launches;
      }
    }
  }
  B200_CHECK(prof_mark(ST_BLUR + 1));
  // K5 orientation + descriptors
  {
    if (!(experimental_mask() & EXP_ORIENT2))
      k_orient_desc<<<dim3((cap + OD_WARPS - 1) / OD_WARPS, F), OD_WARPS * 32, 0, stream>>>(
          ltab, otab, rawv, blurv, d_pattern, d_sel, d_selcnt, sel_per_frame, d_kps, d_desc, d_n, cap);
    else
      k_orient_desc2<<<dim3((cap + OD_WARPS * OD_KPW - 1) / (OD_WARPS * OD_KPW), F), OD_WARPS * 32, 0, stream>>>(
          ltab, otab, rawv, blurv, d_patf, d_sel, d_selcnt, sel_per_frame, d_kps, d_desc, d_n, cap);
  }
  ++launches;
  B200_CHECK(prof_mark(ST_ORIENT_DESC + 1));
  B200_CUDA(cudaGetLastError());
  lastF = f0 + F;
  have_results = true;
  return B200ORB_OK;
}

int orbx::prof_begin(int frames) {
  if (!profile) return B200ORB_OK;
  StageEvents se;
  memset(&se, 0, sizeof(se));
  se.frames = frames;
  prof_runs.push_back(se);
  return prof_mark(0);
}

int orbx::prof_mark(int boundary) {
  if (!profile || prof_runs.empty()) return B200ORB_OK;
  StageEvents& se = prof_runs.back();
  B200_CUDA(cudaEventCreate(&se.ev[boundary]));
  B200_CUDA(cudaEventRecord(se.ev[boundary], stream));
  se.used[boundary] = true;
  return B200ORB_OK;
}

int orbx::ensure_stage(size_t bytes) {
  if (bytes <= stage_bytes) return B200ORB_OK;
  if (h_stage) cudaFreeHost(h_stage);
  h_stage = nullptr; stage_bytes = 0;
  B200_CUDA(cudaHostAlloc(&h_stage, bytes, cudaHostAllocDefault));
  stage_bytes = bytes;
  return B200ORB_OK;
}

int orbx::ensure_tmp(size_t bytes) {
  if (bytes <= tmp_bytes) return B200ORB_OK;
  if (d_tmp) cudaFree(d_tmp);
  d_tmp = nullptr; tmp_bytes = 0;
  B200_CUDA(cudaMalloc(&d_tmp, bytes));
  tmp_bytes = bytes;
  return B200ORB_OK;
}

// =====================================================================================================
// C-ABI
// =====================================================================================================
extern "C" {

int b200orb_experimental(void) { return experimental_mask(); }
int b200orb_get_tuning(int* exp_mask, int* fast_wpc, int* qt_minblocks) {
  const Tuning& t = tuning();
  if (exp_mask) *exp_mask = t.exp_mask;
  if (fast_wpc) *fast_wpc = t.fast_wpc;
  if (qt_minblocks) *qt_minblocks = t.qt_minb;
  return B200ORB_OK;
}
int b200orb_set_tuning(int exp_mask, int fast_wpc, int qt_minblocks) {   // negative = keep
  if ((fast_wpc >= 0 && !valid_wpc(fast_wpc)) || (qt_minblocks >= 0 && !valid_minb(qt_minblocks))) {
    set_error("bad tuning value");
    return B200ORB_EINVAL;
  }
  Tuning& t = tuning();
  if (exp_mask >= 0) t.exp_mask = exp_mask;
  if (fast_wpc >= 0) t.fast_wpc = fast_wpc;
  if (qt_minblocks >= 0) t.qt_minb = qt_minblocks;
  return B200ORB_OK;
}


const char* b200orb_last_error(void) { return g_err; }
const char* b200orb_version(void) { return "b200orb 0.1 (sm_100a)"; }
int b200orb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int orbx_create(const OrbxParams* p, int device, orbx_t** out) {
  if (!p || !out) { set_error("null argument"); return B200ORB_EINVAL; }
  *out = nullptr;
  if (p->nlevels < 1 || p->nlevels > MAX_LEVELS || p->nfeatures < 0 || !(p->scale_factor > 1.0f) ||
      p->ini_th_fast < p->min_th_fast || p->min_th_fast < 0 || p->ini_th_fast > 254) {
    set_error("bad OrbxParams (nlevels 1..%d, scale_factor > 1, 0 <= minTh <= iniTh <= 254)", MAX_LEVELS);
    return B200ORB_EINVAL;
  }
  B200_CHECK(check_device(device));
  DeviceGuard g(device);
  orbx* h = new (std::nothrow) orbx();
  if (!h) { set_error("out of host memory"); return B200ORB_EINVAL; }
  int rc = h->init(*p, device);
  if (rc != B200ORB_OK) { delete h; return rc; }
  *out = h;
  return B200ORB_OK;
}

void orbx_destroy(orbx_t* h) { delete h; }

int orbx_max_keypoints(const orbx_t* h) {
  if (!h) return 0;
  // capacity independent of the geometry: sum over levels of max(N_l + 3, 4*nIni); nIni <= 4 for any image the
  // reference accepts in practice, so 16 is the floor per level
  int s = 0;
  for (int l = 0; l < h->prm.nlevels; ++l) s += std::max(h->nfeat[l] + 3, 16);
  return std::max(s, h->cap);
}

static int extract_common(orbx* h, const uint8_t* gray, bool on_device, int nframes, int rows, int cols,
                          size_t stride, size_t frame_stride) {
  if (!gray) { set_error("null image"); return B200ORB_EINVAL; }
  if (stride < (size_t)cols || (nframes > 1 && frame_stride < stride * (size_t)(rows - 1) + cols)) {
    set_error("bad stride");
    return B200ORB_EINVAL;
  }
  B200_CHECK(h->ensure_geometry(rows, cols, nframes));
  if (on_device) {
    if (stride > 0x7fffffff) { set_error("stride too large"); return B200ORB_EINVAL; }
    return h->run(gray, (int)stride, frame_stride, nframes);
  }
  // host images: stage through the internal level-0 planes (pitch = cols rounded up to 16)
  for (int f = 0; f < nframes; ++f)
    B200_CUDA(cudaMemcpy2DAsync(h->d_raw + (size_t)f * h->frame_bytes, h->lpitch[0], gray + (size_t)f * frame_stride,
                                stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
  return h->run(h->d_raw, h->lpitch[0], h->frame_bytes, nframes);
}

int orbx_extract_batch(orbx_t* h, const uint8_t* gray, int nframes, int rows, int cols, size_t stride,
                       size_t frame_stride, OrbxKeyPoint* kps, uint8_t* desc, int cap, int* n_out) {
  if (!h || !n_out) { set_error("null argument"); return B200ORB_EINVAL; }
  if (nframes <= 0) return B200ORB_OK;
  if (rows <= 0 || cols <= 0) {   // :1055 empty image -> outputs untouched
    for (int f = 0; f < nframes; ++f) n_out[f] = 0;
    return B200ORB_OK;
  }
  if (!kps || !desc || cap <= 0) { set_error("null/empty output buffers"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  B200_CHECK(extract_common(h, gray, false, nframes, rows, cols, stride, frame_stride));
  const size_t kb = sizeof(OrbxKeyPoint) * h->cap, db = (size_t)32 * h->cap;
  B200_CHECK(h->ensure_stage((kb + db) * nframes + sizeof(int) * nframes));
  char* hs = (char*)h->h_stage;
  OrbxKeyPoint* hk = (OrbxKeyPoint*)hs;
  uint8_t* hd = (uint8_t*)(hs + kb * nframes);
  int* hn = (int*)(hs + (kb + db) * nframes);
  B200_CUDA(cudaMemcpyAsync(hn, h->d_n, sizeof(int) * nframes, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(hk, h->d_kps, kb * nframes, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(hd, h->d_desc, db * nframes, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  int rc = B200ORB_OK;
  for (int f = 0; f < nframes; ++f) {
    const int n = hn[f];
    n_out[f] = n;
    if (n > cap) { set_error("frame %d produced %d keypoints > cap %d", f, n, cap); rc = B200ORB_ECAP; continue; }
    memcpy(kps + (size_t)f * cap, hk + (size_t)f * h->cap, sizeof(OrbxKeyPoint) * n);
    memcpy(desc + (size_t)f * cap * 32, hd + (size_t)f * h->cap * 32, (size_t)32 * n);
  }
  return rc;
}

int orbx_extract(orbx_t* h, const uint8_t* gray, int rows, int cols, size_t stride, OrbxKeyPoint* kps,
                 uint8_t* desc, int cap, int* n_out) {
  return orbx_extract_batch(h, gray, 1, rows, cols, stride, stride * (size_t)(rows > 0 ? rows : 0), kps, desc, cap,
                            n_out);
}

int orbx_extract_batch_device(orbx_t* h, const uint8_t* d_gray, int nframes, int rows, int cols, size_t stride,
                              size_t frame_stride) {
  if (!h) { set_error("null handle"); return B200ORB_EINVAL; }
  if (nframes <= 0 || rows <= 0 || cols <= 0) { set_error("empty batch"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  return extract_common(h, d_gray, true, nframes, rows, cols, stride, frame_stride);
}

int orbx_device_results(orbx_t* h, const OrbxKeyPoint** d_kps, const uint8_t** d_desc, const int32_t** d_counts,
                        int* cap) {
  if (!h || !h->have_results) { set_error("no results"); return B200ORB_EINVAL; }
  if (d_kps) *d_kps = h->d_kps;
  if (d_desc) *d_desc = h->d_desc;
  if (d_counts) *d_counts = h->d_n;
  if (cap) *cap = h->cap;
  return B200ORB_OK;
}

int orbx_sync(orbx_t* h) {
  if (!h) return B200ORB_EINVAL;
  DeviceGuard g(h->device);
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}
void* orbx_stream(orbx_t* h) { return h ? (void*)h->stream : nullptr; }

int orbx_level_dims(const orbx_t* h, int level, int* rows, int* cols) {
  if (!h || level < 0 || level >= h->prm.nlevels || h->rows == 0) { set_error("bad level / no geometry yet"); return B200ORB_EINVAL; }
  if (rows) *rows = h->lh[level];
  if (cols) *cols = h->lw[level];
  return B200ORB_OK;
}

int orbx_get_level(orbx_t* h, int frame, int level, int bordered, uint8_t* dst, size_t dst_stride) {
  if (!h || !dst || !h->have_results || level < 0 || level >= h->prm.nlevels || frame < 0 || frame >= h->lastF) {
    set_error("bad argument / no extract yet");
    return B200ORB_EINVAL;
  }
  DeviceGuard g(h->device);
  const int w = h->lw[level], hh = h->lh[level];
  const uint8_t* src = h->rawv.p[level] + (size_t)frame * h->rawv.fstride[level];
  if (!bordered) {
    if (dst_stride < (size_t)w) { set_error("dst_stride too small"); return B200ORB_EINVAL; }
    B200_CUDA(cudaMemcpy2DAsync(dst, dst_stride, src, h->rawv.pitch[level], w, hh, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    return B200ORB_OK;
  }
  const int B = EDGE_THRESHOLD, bw = w + 2 * B, bh = hh + 2 * B;
  if (dst_stride < (size_t)bw) { set_error("dst_stride too small"); return B200ORB_EINVAL; }
  B200_CHECK(h->ensure_tmp((size_t)bw * bh));
  dim3 blk(32, 8), grd((bw + 31) / 32, (bh + 7) / 8);
  k_border_copy<<<grd, blk, 0, h->stream>>>(src, h->rawv.pitch[level], w, hh, (uint8_t*)h->d_tmp, bw, B);
  ++h->launches;
  B200_CUDA(cudaMemcpy2DAsync(dst, dst_stride, h->d_tmp, bw, bw, bh, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int orbx_scale_tables(const orbx_t* h, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2,
                      int* features_per_level) {
  if (!h) return B200ORB_EINVAL;
  for (int i = 0; i < h->prm.nlevels; ++i) {
    if (sf) sf[i] = h->sf[i];
    if (inv_sf) inv_sf[i] = h->invsf[i];
    if (sigma2) sigma2[i] = h->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = h->invsigma2[i];
    if (features_per_level) features_per_level[i] = h->nfeat[i];
  }
  return B200ORB_OK;
}

int orbx_candidates_per_level(orbx_t* h, int frame, int* counts) {
  if (!h || !counts || !h->have_results || frame < 0 || frame >= h->lastF) { set_error("bad argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  B200_CUDA(cudaMemcpyAsync(counts, h->d_candcnt + (size_t)frame * h->prm.nlevels, sizeof(int) * h->prm.nlevels,
                            cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

long long orbx_launch_count(const orbx_t* h) { return h ? h->launches : 0; }

int orbx_profile_enable(orbx_t* h, int on) {
  if (!h) return B200ORB_EINVAL;
  h->profile = on != 0;
  return B200ORB_OK;
}

// Sum of per-stage device time (ms) and frames over every run() since the last read; synchronises the stream.
int orbx_profile_read(orbx_t* h, float* ms /*B200ORB_NUM_STAGES*/, long long* frames, int* runs) {
  if (!h || !ms) return B200ORB_EINVAL;
  DeviceGuard g(h->device);
  B200_CUDA(cudaStreamSynchronize(h->stream));
  for (int s = 0; s < ST_COUNT; ++s) ms[s] = 0.f;
  long long fr = 0;
  for (StageEvents& se : h->prof_runs) {
    fr += se.frames;
    int prev = 0;
    for (int b = 1; b <= ST_COUNT; ++b) {
      if (!se.used[b]) continue;
      float t = 0.f;
      if (se.used[prev]) cudaEventElapsedTime(&t, se.ev[prev], se.ev[b]);
      ms[b - 1] += t;
      prev = b;
    }
    for (int b = 0; b <= ST_COUNT; ++b) if (se.used[b]) cudaEventDestroy(se.ev[b]);
  }
  if (frames) *frames = fr;
  if (runs) *runs = (int)h->prof_runs.size();
  h->prof_runs.clear();
  return B200ORB_OK;
}

}  // extern "C"
