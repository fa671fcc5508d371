// oracle/match_ref.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// CPU restatement of the reference's Hamming searches over flat arrays (the C-ABI views of
// include/b200orb.h): src/ORBmatcher.cc (SearchByProjection x2, SearchByBoW, ComputeThreeMaxima,
// DescriptorDistance) and the Frame helpers they call, src/Frame.cc (AssignFeaturesToGrid, PosInGrid,
// GetFeaturesInArea, ComputeStereoFromRGBD, UnprojectStereo).
//
// Third-party arithmetic restated (parity unpinned): cv::Mat float products.  OpenCV's
// small-matrix gemm path (len<=4, no transpose flags) accumulates a row in float left-to-right and adds the
// "+C" term in double: d = (float)((double)(a0*b0 + a1*b1 + a2*b2) + (double)c).  Products with a
// transposed operand (-Rcw.t()*tcw) go through the generic kernel that accumulates in double.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/b200orb.h"

namespace {

const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;   // src/ORBmatcher.cc:39-41
const int GRID_COLS = 64, GRID_ROWS = 48;                  // include/Frame.h:25-26

// DescriptorDistance, src/ORBmatcher.cc:1968-1984
inline int desc_dist(const uint8_t* a, const uint8_t* b) {
  const uint32_t* pa = (const uint32_t*)a;
  const uint32_t* pb = (const uint32_t*)b;
  int dist = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

inline float gemm3(const float* a, float b0, float b1, float b2, float c) {
  float t = a[0] * b0 + a[1] * b1 + a[2] * b2;
  return (float)((double)t + (double)c);
}

struct Grid {   // Frame::mGrid, src/Frame.cc:319-334
  std::vector<int> cell[GRID_COLS][GRID_ROWS];
  float minX, minY, invW, invH;
  const OrbmFrame* F;
  void build(const OrbmFrame* f) {
    F = f;
    minX = f->min_x; minY = f->min_y;
    invW = static_cast<float>(GRID_COLS) / static_cast<float>(f->max_x - f->min_x);   // src/Frame.cc:221-222
    invH = static_cast<float>(GRID_ROWS) / static_cast<float>(f->max_y - f->min_y);
    for (int i = 0; i < f->n; ++i) {
      int px = (int)std::round((f->x[i] - minX) * invW);   // PosInGrid :522-531 (round, not floor)
      int py = (int)std::round((f->y[i] - minY) * invH);
      if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
      cell[px][py].push_back(i);
    }
  }
  // GetFeaturesInArea :465-518
  void query(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
    out.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
      for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
        for (int idx : cell[ix][iy]) {
          if (bCheckLevels) {
            if (F->octave[idx] < minLevel) continue;
            if (maxLevel >= 0 && F->octave[idx] > maxLevel) continue;
          }
          const float dx = F->x[idx] - x, dy = F->y[idx] - y;
          if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(idx);
        }
  }
};

// ComputeThreeMaxima, src/ORBmatcher.cc:1912-1957
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; ++i) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

inline int rot_bin(float a1, float a2) {   // :1685-1690
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

}  // namespace

extern "C" {

int match_ref_hamming(const uint8_t* a, const uint8_t* b) { return desc_dist(a, b); }

// SearchByProjection(Frame&, const Frame&, th, bMono), src/ORBmatcher.cc:1578-1724
int match_ref_projection_last(const OrbmFrame* cur, const OrbmLast* last, float th, int mono, float nnratio,
                              int check_ori, int32_t* cur2last, int* nmatches_out) {
  (void)nnratio;   // this overload never uses mfNNratio
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  Grid* grid = new Grid();
  grid->build(cur);
  const float* T = cur->Tcw;
  const float Rcw[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  const float tcw[3] = {T[3], T[7], T[11]};
  float twc[3];
  for (int i = 0; i < 3; ++i) {   // -Rcw.t()*tcw : generic gemm, double accumulation, alpha = -1
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)Rcw[k * 3 + i] * (double)tcw[k];
    twc[i] = (float)(-s);
  }
  const float* Tl = last->Tcw;
  const float Rlw2[3] = {Tl[8], Tl[9], Tl[10]};
  const float tlc2 = gemm3(Rlw2, twc[0], twc[1], twc[2], Tl[11]);   // tlc = Rlw*twc + tlw, only z is used
  const bool bForward = tlc2 > cur->b && !mono;
  const bool bBackward = -tlc2 > cur->b && !mono;
  // pointer state of CurrentFrame.mvpMapPoints: -1 NULL, -2 pre-existing, >=0 index into last
  std::vector<int> state(cur->n, -1);
  std::vector<int> state_obs(cur->n, 0);
  for (int j = 0; j < cur->n; ++j)
    if (cur->mp_obs && cur->mp_obs[j] >= 0) { state[j] = -2; state_obs[j] = cur->mp_obs[j]; }
  std::vector<int> cand;
  for (int i = 0; i < last->n; ++i) {
    if (!last->valid[i]) continue;
    const float* X = last->xw + 3 * i;
    const float xc = gemm3(Rcw + 0, X[0], X[1], X[2], tcw[0]);
    const float yc = gemm3(Rcw + 3, X[0], X[1], X[2], tcw[1]);
    const float zc = gemm3(Rcw + 6, X[0], X[1], X[2], tcw[2]);
    const float invzc = (float)(1.0 / zc);
    if (invzc < 0) continue;
    float u = cur->fx * xc * invzc + cur->cx;
    float v = cur->fy * yc * invzc + cur->cy;
    if (std::isnan(u) || std::isnan(v)) continue;   // reference: UB (zc == 0 with xc == 0); documented deviation
    if (u < cur->min_x || u > cur->max_x) continue;
    if (v < cur->min_y || v > cur->max_y) continue;
    const int nLastOctave = last->octave[i];
    const float radius = th * cur->scale_factors[nLastOctave];
    if (bForward) grid->query(u, v, radius, nLastOctave, -1, cand);
    else if (bBackward) grid->query(u, v, radius, 0, nLastOctave, cand);
    else grid->query(u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
    if (cand.empty()) continue;
    const uint8_t* dMP = last->mp_desc + 32 * (size_t)i;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (state[i2] != -1 && state_obs[i2] > 0) continue;
      if (cur->uright[i2] > 0) {
        const float ur = u - cur->bf * invzc;
        const float er = std::fabs(ur - cur->uright[i2]);
        if (er > radius) continue;
      }
      const int dist = desc_dist(dMP, cur->desc + 32 * (size_t)i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= TH_HIGH) {
      state[bestIdx2] = i;
      state_obs[bestIdx2] = last->mp_obs ? last->mp_obs[i] : 0;
      nmatches++;
      if (check_ori) rotHist[rot_bin(last->angle[i], cur->angle[bestIdx2])].push_back(bestIdx2);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int idx : rotHist[i]) { state[idx] = -1; nmatches--; }
  }
  for (int j = 0; j < cur->n; ++j) cur2last[j] = state[j];
  *nmatches_out = nmatches;
  delete grid;
  return 0;
}

// SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:63-156
int match_ref_projection_points(const OrbmFrame* F, const OrbmTrackPoints* pts, float th, float nnratio,
                                int32_t* f2pt, int* nmatches_out) {
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  Grid* grid = new Grid();
  grid->build(F);
  std::vector<int> state(F->n, -1), state_obs(F->n, 0);
  for (int j = 0; j < F->n; ++j)
    if (F->mp_obs && F->mp_obs[j] >= 0) { state[j] = -2; state_obs[j] = F->mp_obs[j]; }
  std::vector<int> cand;
  for (int iMP = 0; iMP < pts->n; ++iMP) {
    if (!pts->track_in_view[iMP]) continue;
    const int nPredictedLevel = pts->scale_level[iMP];
    float r = (pts->view_cos[iMP] > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos :159-165 (compare in double)
    if (bFactor) r *= th;
    const float rs = r * F->scale_factors[nPredictedLevel];
    grid->query(pts->proj_x[iMP], pts->proj_y[iMP], rs, nPredictedLevel - 1, nPredictedLevel, cand);
    if (cand.empty()) continue;
    const uint8_t* dMP = pts->mp_desc + 32 * (size_t)iMP;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cand) {
      if (state[idx] != -1 && state_obs[idx] > 0) continue;
      if (F->uright[idx] > 0) {
        const float er = std::fabs(pts->proj_xr[iMP] - F->uright[idx]);
        if (er > rs) continue;
      }
      const int dist = desc_dist(dMP, F->desc + 32 * (size_t)idx);
      if (dist < bestDist) {
        bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F->octave[idx]; bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = F->octave[idx]; bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      state[bestIdx] = iMP;
      state_obs[bestIdx] = pts->mp_obs ? pts->mp_obs[iMP] : 1;
      nmatches++;
    }
  }
  for (int j = 0; j < F->n; ++j) f2pt[j] = state[j];
  *nmatches_out = nmatches;
  delete grid;
  return 0;
}

// Generic guided search behind orbm_search_projected: the shared tail of the projection overloads whose query geometry
// the caller computes -- candidate loop + "already matched" rule + rotation histogram, e.g. src/ORBmatcher.cc:1804-1897
// (relocalisation: claim_rule 1, bestDist <= ORBdist) or :1645-1721 (claim_rule 0).
int match_ref_projected(const OrbmFrame* cur, const OrbmQueries* q, int max_dist, int claim_rule, int check_ori,
                        int32_t* cur2q, int* nmatches_out) {
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  Grid* grid = new Grid();
  grid->build(cur);
  std::vector<int> state(cur->n, -1), state_obs(cur->n, 0);
  for (int j = 0; j < cur->n; ++j)
    if (cur->mp_obs && cur->mp_obs[j] >= 0) { state[j] = -2; state_obs[j] = cur->mp_obs[j]; }
  std::vector<int> cand;
  for (int i = 0; i < q->n; ++i) {
    if (!q->valid[i]) continue;
    const float u = q->u[i], v = q->v[i], radius = q->radius[i];
    if (std::isnan(u) || std::isnan(v)) continue;
    grid->query(u, v, radius, q->min_level[i], q->max_level[i], cand);
    if (cand.empty()) continue;
    const uint8_t* dMP = q->desc + 32 * (size_t)i;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (claim_rule) { if (state[i2] != -1) continue; }
      else if (state[i2] != -1 && state_obs[i2] > 0) continue;
      if (q->uright && cur->uright[i2] > 0) {
        const float er = std::fabs(q->uright[i] - cur->uright[i2]);
        if (er > radius) continue;
      }
      const int dist = desc_dist(dMP, cur->desc + 32 * (size_t)i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= max_dist) {
      state[bestIdx2] = i;
      state_obs[bestIdx2] = q->obs ? q->obs[i] : 1;
      nmatches++;
      if (check_ori) rotHist[rot_bin(q->angle[i], cur->angle[bestIdx2])].push_back(bestIdx2);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int idx : rotHist[i]) { state[idx] = -1; nmatches--; }
  }
  for (int j = 0; j < cur->n; ++j) cur2q[j] = state[j];
  *nmatches_out = nmatches;
  delete grid;
  return 0;
}

// SearchForInitialization(Frame& F1, Frame& F2, vector<Point2f>& vbPrevMatched, vector<int>& vnMatches12, int windowSize),
// src/ORBmatcher.cc:523-660 (monocular initialisation; not behind the C-ABI yet -- the oracle is here so that the
// kernel of the next round has its checker).  F1 supplies keypoints (octave, angle) and descriptors, F2 additionally
// its grid; prev_xy[n1][2] is vbPrevMatched (in/out), matches12[n1] the result.  Differences from the projection
// searches: level 0 only, a closer match STEALS an already matched F2 keypoint (vMatchedDistance), ratio test against
// the second best, and the orientation prune only clears entries that are still matched.
int match_ref_initialization(const OrbmFrame* F1, const OrbmFrame* F2, float* prev_xy, int window, float nnratio,
                             int check_ori, int32_t* matches12, int* nmatches_out) {
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  Grid* grid = new Grid();
  grid->build(F2);
  std::vector<int> vnMatches12(F1->n, -1), vnMatches21(F2->n, -1), vMatchedDistance(F2->n, 2147483647);
  std::vector<int> cand;
  for (int i1 = 0; i1 < F1->n; ++i1) {
    const int level1 = F1->octave[i1];
    if (level1 > 0) continue;
    grid->query(prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window, level1, level1, cand);
    if (cand.empty()) continue;
    const uint8_t* d1 = F1->desc + 32 * (size_t)i1;
    int bestDist = 2147483647, bestDist2 = 2147483647, bestIdx2 = -1;
    for (int i2 : cand) {
      const int dist = desc_dist(d1, F2->desc + 32 * (size_t)i2);
      if (vMatchedDistance[i2] <= dist) continue;
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
      else if (dist < bestDist2) bestDist2 = dist;
    }
    if (bestDist <= TH_LOW) {
      if (bestDist < (float)bestDist2 * nnratio) {
        if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
        vnMatches12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        nmatches++;
        if (check_ori) rotHist[rot_bin(F1->angle[i1], F2->angle[bestIdx2])].push_back(i1);
      }
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
    }
  }
  for (int i1 = 0; i1 < F1->n; ++i1) {
    matches12[i1] = vnMatches12[i1];
    if (vnMatches12[i1] >= 0) { prev_xy[2 * i1] = F2->x[vnMatches12[i1]]; prev_xy[2 * i1 + 1] = F2->y[vnMatches12[i1]]; }
  }
  *nmatches_out = nmatches;
  delete grid;
  return 0;
}

// The candidate loop that Fuse (src/ORBmatcher.cc:1106-1155, gate 1), Fuse with Sim3 (:1268-1287) and both directions of
// SearchBySim3 (:1421-1444, :1494-1517) share: KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:659-698: no level filter
// inside), level range [min_level, max_level] = [nPredictedLevel-1, nPredictedLevel] in the loop, the Fuse reprojection
// gate, first minimum of the Hamming distance.  No claim state.  best_dist = INT_MAX when there is no candidate.
int match_ref_best(const OrbmFrame* kf, const OrbmQueries* q, int gate, const float* inv_level_sigma2, int32_t* best_idx,
                   int32_t* best_dist) {
  Grid* grid = new Grid();
  grid->build(kf);
  std::vector<int> cand;
  for (int i = 0; i < q->n; ++i) {
    best_idx[i] = -1; best_dist[i] = 2147483647;
    if (!q->valid[i]) continue;
    const float u = q->u[i], v = q->v[i];
    if (std::isnan(u) || std::isnan(v)) continue;
    grid->query(u, v, q->radius[i], -1, -1, cand);
    const uint8_t* dMP = q->desc + 32 * (size_t)i;
    int bestDist = 2147483647, bestIdx = -1;
    for (int idx : cand) {
      const int kpLevel = kf->octave[idx];
      if (kpLevel < q->min_level[i] || kpLevel > q->max_level[i]) continue;
      if (gate == 1) {
        const float kpx = kf->x[idx], kpy = kf->y[idx];
        if (kf->uright[idx] >= 0) {
          const float ur = q->uright[i];
          const float kpr = kf->uright[idx];
          const float ex = u - kpx, ey = v - kpy, er = ur - kpr;
          const float e2 = ex * ex + ey * ey + er * er;
          if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
        } else {
          const float ex = u - kpx, ey = v - kpy;
          const float e2 = ex * ex + ey * ey;
          if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
        }
      }
      const int dist = desc_dist(dMP, kf->desc + 32 * (size_t)idx);
      if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    best_idx[i] = bestIdx; best_dist[i] = bestDist;
  }
  delete grid;
  return 0;
}

// SearchForTriangulation(KeyFrame*, KeyFrame*, cv::Mat F12, vector<pair<size_t,size_t>>&, bOnlyStereo),
// src/ORBmatcher.cc:827-1019 (+ CheckDistEpipolarLine :175-194).  (ex, ey) = epipole of camera 1 in image 2 (:835-839).
// The reference never sets vbMatched2, so several pKF1 keypoints may end on the same pKF2 keypoint.
int match_ref_triangulation(const OrbmTriKF* k1, const OrbmTriKF* k2, const float* F12, float ex, float ey,
                            const float* sf2, const float* sigma2_2, int only_stereo, int check_ori, int32_t* matches12,
                            int* nmatches_out) {
  int nmatches = 0;
  std::vector<int> vMatches12(k1->n, -1);
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < k1->n_nodes && b < k2->n_nodes) {
    if (k1->node_ids[a] == k2->node_ids[b]) {
      for (int i1 = k1->node_off[a]; i1 < k1->node_off[a + 1]; ++i1) {
        const int idx1 = (int)k1->idx[i1];
        if (k1->has_mp && k1->has_mp[idx1]) continue;
        const bool bStereo1 = k1->uright[idx1] >= 0;
        if (only_stereo && !bStereo1) continue;
        const uint8_t* d1 = k1->desc + 32 * (size_t)idx1;
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int i2 = k2->node_off[b]; i2 < k2->node_off[b + 1]; ++i2) {
          const int idx2 = (int)k2->idx[i2];
          if (k2->has_mp && k2->has_mp[idx2]) continue;
          const bool bStereo2 = k2->uright[idx2] >= 0;
          if (only_stereo && !bStereo2) continue;
          const int dist = desc_dist(d1, k2->desc + 32 * (size_t)idx2);
          if (dist > TH_LOW || dist > bestDist) continue;
          if (!bStereo1 && !bStereo2) {
            const float distex = ex - k2->x[idx2], distey = ey - k2->y[idx2];
            if (distex * distex + distey * distey < 100 * sf2[k2->octave[idx2]]) continue;
          }
          // CheckDistEpipolarLine
          const float x1 = k1->x[idx1], y1 = k1->y[idx1];
          const float la = x1 * F12[0] + y1 * F12[3] + F12[6];
          const float lb = x1 * F12[1] + y1 * F12[4] + F12[7];
          const float lc = x1 * F12[2] + y1 * F12[5] + F12[8];
          const float num = la * k2->x[idx2] + lb * k2->y[idx2] + lc;
          const float den = la * la + lb * lb;
          if (den == 0) continue;
          const float dsqr = num * num / den;
          if (dsqr < 3.84 * sigma2_2[k2->octave[idx2]]) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
          vMatches12[idx1] = bestIdx2;
          nmatches++;
          if (check_ori) rotHist[rot_bin(k1->angle[idx1], k2->angle[bestIdx2])].push_back(idx1);
        }
      }
      ++a; ++b;
    } else if (k1->node_ids[a] < k2->node_ids[b]) {
      a = (int)(std::lower_bound(k1->node_ids, k1->node_ids + k1->n_nodes, k2->node_ids[b]) - k1->node_ids);
    } else {
      b = (int)(std::lower_bound(k2->node_ids, k2->node_ids + k2->n_nodes, k1->node_ids[a]) - k2->node_ids);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) { vMatches12[idx1] = -1; nmatches--; }
    }
  }
  for (int i = 0; i < k1->n; ++i) matches12[i] = vMatches12[i];
  *nmatches_out = nmatches;
  return 0;
}

// Frame::isInFrustum (src/Frame.cc:387-451) + MapPoint::PredictScale (src/MapPoint.cc:463-478) for n MapPoints.
// min_dist / max_dist = the raw members mfMinDistance / mfMaxDistance.  Entries that are not in view keep 0.
int match_ref_is_in_frustum(const OrbmFrame* f, int n, const float* xw, const float* normal, const float* min_dist,
                            const float* max_dist, float viewingCosLimit, float log_scale_factor, uint8_t* in_view,
                            float* proj_x, float* proj_y, float* proj_xr, int32_t* scale_level, float* view_cos) {
  const float* T = f->Tcw;
  float Ow[3];
  for (int a = 0; a < 3; ++a) {   // mOw = -mRwc*mtcw (src/Frame.cc:373): evaluated transpose -> small-matrix gemm, float sum
    const float t = T[0 * 4 + a] * T[3] + T[1 * 4 + a] * T[7] + T[2 * 4 + a] * T[11];
    Ow[a] = -t;
  }
  for (int i = 0; i < n; ++i) {
    in_view[i] = 0; proj_x[i] = proj_y[i] = proj_xr[i] = view_cos[i] = 0.f; scale_level[i] = 0;
    const float* P = xw + 3 * i;
    const float PcX = gemm3(T + 0, P[0], P[1], P[2], T[3]);
    const float PcY = gemm3(T + 4, P[0], P[1], P[2], T[7]);
    const float PcZ = gemm3(T + 8, P[0], P[1], P[2], T[11]);
    if (PcZ < 0.0f) continue;
    const float invz = 1.0f / PcZ;
    const float u = f->fx * PcX * invz + f->cx;
    const float v = f->fy * PcY * invz + f->cy;
    if (u < f->min_x || u > f->max_x) continue;
    if (v < f->min_y || v > f->max_y) continue;
    const float maxDistance = 1.2f * max_dist[i];
    const float minDistance = 0.8f * min_dist[i];
    const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
    double n2 = 0, dot = 0;
    for (int a = 0; a < 3; ++a) { n2 += (double)PO[a] * (double)PO[a]; dot += (double)PO[a] * (double)normal[3 * i + a]; }
    const float dist = (float)std::sqrt(n2);                 // cv::norm: double accumulation
    if (dist < minDistance || dist > maxDistance) continue;
    const float viewCos = (float)(dot / dist);               // Mat::dot returns double
    if (viewCos < viewingCosLimit) continue;
    const float ratio = max_dist[i] / dist;
    int nScale = (int)std::ceil(std::log(ratio) / log_scale_factor);   // float log, float division, float ceil
    if (nScale < 0) nScale = 0;
    else if (nScale >= f->nlevels) nScale = f->nlevels - 1;
    in_view[i] = 1;
    proj_x[i] = u; proj_y[i] = v; proj_xr[i] = u - f->bf * invz;
    scale_level[i] = nScale; view_cos[i] = viewCos;
  }
  return 0;
}

// SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), src/ORBmatcher.cc:217-363
int match_ref_bow(const OrbmBow* kf, const OrbmBow* f, float nnratio, int check_ori, int32_t* f2kf,
                  int* nmatches_out) {
  std::vector<int> matches(f->n, -1);
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < kf->n_nodes && b < f->n_nodes) {
    if (kf->node_ids[a] == f->node_ids[b]) {
      for (int iKF = kf->node_off[a]; iKF < kf->node_off[a + 1]; ++iKF) {
        const unsigned realIdxKF = kf->idx[iKF];
        if (kf->valid && !kf->valid[realIdxKF]) continue;
        const uint8_t* dKF = kf->desc + 32 * (size_t)realIdxKF;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int iF = f->node_off[b]; iF < f->node_off[b + 1]; ++iF) {
          const unsigned realIdxF = f->idx[iF];
          if (matches[realIdxF] >= 0) continue;
          const int dist = desc_dist(dKF, f->desc + 32 * (size_t)realIdxF);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = (int)realIdxF; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= TH_LOW) {
          if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
            matches[bestIdxF] = (int)realIdxKF;
            if (check_ori) rotHist[rot_bin(kf->angle[realIdxKF], f->angle[bestIdxF])].push_back(bestIdxF);
            nmatches++;
          }
        }
      }
      ++a; ++b;
    } else if (kf->node_ids[a] < f->node_ids[b]) {
      a = (int)(std::lower_bound(kf->node_ids, kf->node_ids + kf->n_nodes, f->node_ids[b]) - kf->node_ids);
    } else {
      b = (int)(std::lower_bound(f->node_ids, f->node_ids + f->n_nodes, kf->node_ids[a]) - f->node_ids);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { matches[idx] = -1; nmatches--; }
    }
  }
  for (int j = 0; j < f->n; ++j) f2kf[j] = matches[j];
  *nmatches_out = nmatches;
  return 0;
}

// SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12), src/ORBmatcher.cc:665-812
int match_ref_bow_kf(const OrbmBow* k1, const OrbmBow* k2, float nnratio, int check_ori, int32_t* matches12,
                     int* nmatches_out) {
  std::vector<int> m12(k1->n, -1);
  std::vector<char> matched2(k2->n, 0);
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < k1->n_nodes && b < k2->n_nodes) {
    if (k1->node_ids[a] == k2->node_ids[b]) {
      for (int i1 = k1->node_off[a]; i1 < k1->node_off[a + 1]; ++i1) {
        const unsigned idx1 = k1->idx[i1];
        if (k1->valid && !k1->valid[idx1]) continue;
        const uint8_t* d1 = k1->desc + 32 * (size_t)idx1;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int i2 = k2->node_off[b]; i2 < k2->node_off[b + 1]; ++i2) {
          const unsigned idx2 = k2->idx[i2];
          if (matched2[idx2] || (k2->valid && !k2->valid[idx2])) continue;
          const int dist = desc_dist(d1, k2->desc + 32 * (size_t)idx2);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = (int)idx2; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 < TH_LOW) {
          if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
            m12[idx1] = bestIdx2;
            matched2[bestIdx2] = 1;
            if (check_ori) rotHist[rot_bin(k1->angle[idx1], k2->angle[bestIdx2])].push_back((int)idx1);
            nmatches++;
          }
        }
      }
      ++a; ++b;
    } else if (k1->node_ids[a] < k2->node_ids[b]) {
      a = (int)(std::lower_bound(k1->node_ids, k1->node_ids + k1->n_nodes, k2->node_ids[b]) - k1->node_ids);
    } else {
      b = (int)(std::lower_bound(k2->node_ids, k2->node_ids + k2->n_nodes, k1->node_ids[a]) - k2->node_ids);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { m12[idx] = -1; nmatches--; }
    }
  }
  for (int j = 0; j < k1->n; ++j) matches12[j] = m12[j];
  *nmatches_out = nmatches;
  return 0;
}

// Frame::ComputeStereoFromRGBD (src/Frame.cc:850-871) + Frame::UnprojectStereo (:879-899) for every keypoint.
// kps: n x (x,y) pairs with stride `kp_stride` floats (OrbxKeyPoint = 7). depth: rows x cols f32.
// Outputs: uright[n], depth_out[n] (-1 where d<=0), xw[n*3] (untouched where d<=0), valid[n].
void frame_ref_stereo_unproject(const float* kps, int kp_stride, int n, const float* depth, int rows, int cols,
                                const float* Tcw, float fx, float fy, float cx, float cy, float bf, float* uright,
                                float* depth_out, float* xw, uint8_t* valid) {
  (void)rows;
  const float invfx = 1.0f / fx, invfy = 1.0f / fy;   // src/Frame.cc:215-216
  // mRwc = mRcw.t(); mOw = -mRwc*mtcw (src/Frame.cc:364-373: the product of the EVALUATED transpose goes through the
  // small-matrix gemm path -- float accumulation, alpha = -1; upstream ORB-SLAM2 writes -mRcw.t()*mtcw, which would
  // accumulate in double; pinned by tests/test_refsrc_cpu.py against the reference's own Frame.cc)
  const float Rcw[9] = {Tcw[0], Tcw[1], Tcw[2], Tcw[4], Tcw[5], Tcw[6], Tcw[8], Tcw[9], Tcw[10]};
  const float tcw[3] = {Tcw[3], Tcw[7], Tcw[11]};
  float Rwc[9], Ow[3];
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) Rwc[i * 3 + k] = Rcw[k * 3 + i];
    const float t = Rwc[i * 3 + 0] * tcw[0] + Rwc[i * 3 + 1] * tcw[1] + Rwc[i * 3 + 2] * tcw[2];
    Ow[i] = -t;
  }
  for (int i = 0; i < n; ++i) {
    const float u = kps[(size_t)i * kp_stride], v = kps[(size_t)i * kp_stride + 1];
    const float d = depth[(size_t)(int)v * cols + (int)u];   // imDepth.at<float>(v,u): float -> int truncation
    uright[i] = -1; depth_out[i] = -1; valid[i] = 0;
    if (d > 0) {
      depth_out[i] = d;
      uright[i] = u - bf / d;
      const float x = (u - cx) * d * invfx, y = (v - cy) * d * invfy;
      xw[3 * i + 0] = gemm3(Rwc + 0, x, y, d, Ow[0]);
      xw[3 * i + 1] = gemm3(Rwc + 3, x, y, d, Ow[1]);
      xw[3 * i + 2] = gemm3(Rwc + 6, x, y, d, Ow[2]);
      valid[i] = 1;
    }
  }
}

}  // extern "C"
