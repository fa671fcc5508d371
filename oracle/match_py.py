"""oracle/match_py.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Second, independent restatement (pure Python + numpy float32 scalars; small cases only) of
ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (reference src/ORBmatcher.cc:1578-1724),
SearchByProjection(Frame&, vector<MapPoint*>&, th) (:63-156), SearchByBoW(KF,F) (:217-363), SearchByBoW(KF,KF) (:665-812)
and the Frame grid helpers (src/Frame.cc:319-334,465-531).  It pins oracle/match_ref.cpp and generates the golden match
vectors under tests/golden/ (tools/make_golden.py).  Reads the same FrameView / LastView objects.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30
GRID_COLS, GRID_ROWS = 64, 48


def c_round(v: float) -> int:
    """C round(): half away from zero."""
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


def hamming(a, b) -> int:
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def gemm3(a, b, c):
    """OpenCV small-matrix gemm row: float accumulation left to right, '+C' in double."""
    t = f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))
    return f32(float(t) + float(c))


class Grid:
    def __init__(self, F):
        self.F = F
        self.minx, self.maxx, self.miny, self.maxy = (f32(v) for v in F.bounds)
        self.invw = f32(f32(GRID_COLS) / f32(self.maxx - self.minx))
        self.invh = f32(f32(GRID_ROWS) / f32(self.maxy - self.miny))
        self.cells = [[[] for _ in range(GRID_ROWS)] for _ in range(GRID_COLS)]
        for i in range(F.n):
            px = c_round(float(f32(f32(F.x[i] - self.minx) * self.invw)))
            py = c_round(float(f32(f32(F.y[i] - self.miny) * self.invh)))
            if px < 0 or px >= GRID_COLS or py < 0 or py >= GRID_ROWS:
                continue
            self.cells[px][py].append(i)

    def query(self, x, y, r, minLevel, maxLevel):
        F = self.F
        out = []
        nMinCellX = max(0, int(math.floor(float(f32(f32(f32(x - self.minx) - r) * self.invw)))))
        if nMinCellX >= GRID_COLS:
            return out
        nMaxCellX = min(GRID_COLS - 1, int(math.ceil(float(f32(f32(f32(x - self.minx) + r) * self.invw)))))
        if nMaxCellX < 0:
            return out
        nMinCellY = max(0, int(math.floor(float(f32(f32(f32(y - self.miny) - r) * self.invh)))))
        if nMinCellY >= GRID_ROWS:
            return out
        nMaxCellY = min(GRID_ROWS - 1, int(math.ceil(float(f32(f32(f32(y - self.miny) + r) * self.invh)))))
        if nMaxCellY < 0:
            return out
        check = (minLevel > 0) or (maxLevel >= 0)
        for ix in range(nMinCellX, nMaxCellX + 1):
            for iy in range(nMinCellY, nMaxCellY + 1):
                for idx in self.cells[ix][iy]:
                    if check:
                        if F.octave[idx] < minLevel:
                            continue
                        if maxLevel >= 0 and F.octave[idx] > maxLevel:
                            continue
                    if abs(f32(F.x[idx] - x)) < r and abs(f32(F.y[idx] - y)) < r:
                        out.append(idx)
        return out


def three_maxima(histo):
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, h in enumerate(histo):
        s = len(h)
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if f32(max2) < f32(f32(0.1) * f32(max1)):
        ind2 = ind3 = -1
    elif f32(max3) < f32(f32(0.1) * f32(max1)):
        ind3 = -1
    return ind1, ind2, ind3


def search_by_projection_last(cur, last, th, mono=False, check_ori=True):
    th = f32(th)
    fx, fy, cx, cy, bf = (f32(v) for v in cur.cam)
    mb = f32(bf / fx)
    T = cur.Tcw.reshape(4, 4)
    Rcw, tcw = T[:3, :3], T[:3, 3]
    twc = [f32(-sum(float(Rcw[k, i]) * float(tcw[k]) for k in range(3))) for i in range(3)]
    Tl = last.Tcw.reshape(4, 4)
    tlc2 = gemm3(Tl[2, :3], twc, Tl[2, 3])
    bForward = (tlc2 > mb) and not mono
    bBackward = (-tlc2 > mb) and not mono
    grid = Grid(cur)
    state = [-1] * cur.n
    obs = [0] * cur.n
    if cur.mp_obs is not None:
        for j in range(cur.n):
            if cur.mp_obs[j] >= 0:
                state[j], obs[j] = -2, int(cur.mp_obs[j])
    rot = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    for i in range(last.n):
        if not last.valid[i]:
            continue
        X = last.xw[i]
        xc = gemm3(Rcw[0], X, tcw[0])
        yc = gemm3(Rcw[1], X, tcw[1])
        zc = gemm3(Rcw[2], X, tcw[2])
        with np.errstate(divide="ignore"):
            invzc = f32(np.float64(1.0) / np.float64(zc))
        if invzc < 0:
            continue
        u = f32(f32(f32(fx * xc) * invzc) + cx)
        v = f32(f32(f32(fy * yc) * invzc) + cy)
        if np.isnan(u) or np.isnan(v):
            continue
        if u < grid.minx or u > grid.maxx or v < grid.miny or v > grid.maxy:
            continue
        octv = int(last.octave[i])
        radius = f32(th * cur.scale_factors[octv])
        if bForward:
            cand = grid.query(u, v, radius, octv, -1)
        elif bBackward:
            cand = grid.query(u, v, radius, 0, octv)
        else:
            cand = grid.query(u, v, radius, octv - 1, octv + 1)
        if not cand:
            continue
        best, bidx = 256, -1
        for i2 in cand:
            if state[i2] != -1 and obs[i2] > 0:
                continue
            if cur.uright[i2] > 0:
                ur = f32(u - f32(bf * invzc))
                if abs(f32(ur - cur.uright[i2])) > radius:
                    continue
            d = hamming(last.mp_desc[i], cur.desc[i2])
            if d < best:
                best, bidx = d, i2
        if best <= TH_HIGH:
            state[bidx] = i
            obs[bidx] = int(last.mp_obs[i]) if last.mp_obs is not None else 0
            nmatches += 1
            if check_ori:
                r = f32(last.angle[i] - cur.angle[bidx])
                if r < 0:
                    r = f32(r + f32(360.0))
                b = c_round(float(f32(r * factor)))
                if b == HISTO_LENGTH:
                    b = 0
                rot[b].append(bidx)
    if check_ori:
        i1, i2_, i3 = three_maxima(rot)
        for b in range(HISTO_LENGTH):
            if b not in (i1, i2_, i3):
                for idx in rot[b]:
                    state[idx] = -1
                    nmatches -= 1
    return nmatches, np.array(state, np.int32)


def _rot_bin(a_query, a_cur, factor):
    r = f32(f32(a_query) - f32(a_cur))
    if r < 0:
        r = f32(r + f32(360.0))
    b = c_round(float(f32(r * factor)))
    return 0 if b == HISTO_LENGTH else b


def search_by_projection_points(F, pts, th, nnratio=0.8):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:63-156) on a
    FrameView + TrackPointsView.  Returns (nmatches, frame->point index vector; -2 = untouched pre-existing point)."""
    th = f32(th)
    nnratio = f32(nnratio)
    grid = Grid(F)
    state = [-1] * F.n
    obs = [0] * F.n
    if F.mp_obs is not None:
        for j in range(F.n):
            if F.mp_obs[j] >= 0:
                state[j], obs[j] = -2, int(F.mp_obs[j])
    b_factor = th != f32(1.0)
    nmatches = 0
    for i in range(pts.n):
        if not pts.track_in_view[i]:
            continue
        lvl = int(pts.scale_level[i])
        r = f32(2.5) if pts.view_cos[i] > f32(0.998) else f32(4.0)     # RadiusByViewingCos :158-164
        if b_factor:
            r = f32(r * th)
        win = f32(r * F.scale_factors[lvl])
        cand = grid.query(f32(pts.proj_x[i]), f32(pts.proj_y[i]), win, lvl - 1, lvl)
        if not cand:
            continue
        best, best_lvl, best2, best_lvl2, bidx = 256, -1, 256, -1, -1
        for idx in cand:
            if state[idx] != -1 and obs[idx] > 0:
                continue
            if F.uright[idx] > 0:
                if abs(f32(f32(pts.proj_xr[i]) - F.uright[idx])) > win:
                    continue
            d = hamming(pts.mp_desc[i], F.desc[idx])
            if d < best:
                best2, best_lvl2 = best, best_lvl
                best, best_lvl, bidx = d, int(F.octave[idx]), idx
            elif d < best2:
                best_lvl2, best2 = int(F.octave[idx]), d
        if best <= TH_HIGH:
            if best_lvl == best_lvl2 and f32(best) > f32(nnratio * f32(best2)):
                continue
            state[bidx] = i
            obs[bidx] = int(pts.mp_obs[i]) if pts.mp_obs is not None else 1
            nmatches += 1
    return nmatches, np.array(state, np.int32)


def _feature_vector(view):
    return {int(view.node_ids[k]): [int(v) for v in view.idx[view.node_off[k]:view.node_off[k + 1]]]
            for k in range(len(view.node_ids))}


def _merge_join(fv1, fv2):
    """The two-iterator walk over the ordered FeatureVector maps (:233-320 / :690-770): yields the index lists of equal
    node ids in ascending node order (lower_bound jumps only skip unequal ids)."""
    for node in sorted(set(fv1) & set(fv2)):
        yield fv1[node], fv2[node]


def search_by_bow(kf, f, nnratio=0.7, check_ori=True):
    """SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:217-363) on two BowViews (kf.valid = the
    keyframe keypoint holds a good MapPoint).  Returns (nmatches, frame keypoint -> keyframe keypoint, -1 = none)."""
    nnratio = f32(nnratio)
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    out = [-1] * f.n
    rot = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    for idx_kf, idx_f in _merge_join(_feature_vector(kf), _feature_vector(f)):
        for rk in idx_kf:
            if kf.valid is not None and not kf.valid[rk]:
                continue
            best1, best2, bidx = 256, 256, -1
            for rf in idx_f:
                if out[rf] != -1:
                    continue
                d = hamming(kf.desc[rk], f.desc[rf])
                if d < best1:
                    best2, best1, bidx = best1, d, rf
                elif d < best2:
                    best2 = d
            if best1 <= TH_LOW and f32(best1) < f32(nnratio * f32(best2)):
                out[bidx] = rk
                if check_ori:
                    rot[_rot_bin(kf.angle[rk], f.angle[bidx], factor)].append(bidx)
                nmatches += 1
    if check_ori:
        keep = three_maxima(rot)
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for j in rot[b]:
                    out[j] = -1
                    nmatches -= 1
    return nmatches, np.array(out, np.int32)


def search_by_bow_kf(kf1, kf2, nnratio=0.75, check_ori=True):
    """SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (src/ORBmatcher.cc:665-812).  Returns (nmatches,
    keyframe-1 keypoint -> keyframe-2 keypoint, -1 = none)."""
    nnratio = f32(nnratio)
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    out = [-1] * kf1.n
    matched2 = [False] * kf2.n
    rot = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    for idx1, idx2 in _merge_join(_feature_vector(kf1), _feature_vector(kf2)):
        for i1 in idx1:
            if kf1.valid is not None and not kf1.valid[i1]:
                continue
            best1, best2, bidx = 256, 256, -1
            for i2 in idx2:
                if matched2[i2] or (kf2.valid is not None and not kf2.valid[i2]):
                    continue
                d = hamming(kf1.desc[i1], kf2.desc[i2])
                if d < best1:
                    best2, best1, bidx = best1, d, i2
                elif d < best2:
                    best2 = d
            if best1 < TH_LOW and f32(best1) < f32(nnratio * f32(best2)):     # strict '<' here (:741), '<=' in (KF,F)
                out[i1] = bidx
                matched2[bidx] = True
                if check_ori:
                    rot[_rot_bin(kf1.angle[i1], kf2.angle[bidx], factor)].append(i1)
                nmatches += 1
    if check_ori:
        keep = three_maxima(rot)
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for j in rot[b]:
                    out[j] = -1
                    nmatches -= 1
    return nmatches, np.array(out, np.int32)


def search_projected(F, q, max_dist, claim_rule=1, check_ori=True):
    """The part the remaining projection overloads share, from the candidate query on, written after the loop bodies of
    SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist) (src/ORBmatcher.cc:1800-1850: rule 1, any MapPoint on the
    candidate blocks it) and of the LAST overload (:1645-1693: rule 0, only MapPoints with Observations() > 0 block;
    stereo gate when a predicted right coordinate is supplied).  q = QueriesView with per-query (u, v, radius, level
    range, descriptor, angle).  Returns (nmatches, keypoint -> query index; -2 = untouched pre-existing point)."""
    grid = Grid(F)
    state = [-1] * F.n
    obs = [0] * F.n
    if F.mp_obs is not None:
        for j in range(F.n):
            if F.mp_obs[j] >= 0:
                state[j], obs[j] = -2, int(F.mp_obs[j])
    rot = [[] for _ in range(HISTO_LENGTH)]
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    nmatches = 0
    for i in range(q.n):
        if not q.valid[i]:
            continue
        u, v, radius = f32(q.u[i]), f32(q.v[i]), f32(q.radius[i])
        if np.isnan(u) or np.isnan(v):
            continue
        cand = grid.query(u, v, radius, int(q.min_level[i]), int(q.max_level[i]))
        if not cand:
            continue
        best, bidx = 256, -1
        for i2 in cand:
            if state[i2] != -1:                       # a MapPoint sits on the candidate
                if claim_rule == 1 or obs[i2] > 0:
                    continue
            if q.uright is not None and F.uright[i2] > 0:
                if abs(f32(f32(q.uright[i]) - F.uright[i2])) > radius:
                    continue
            d = hamming(q.desc[i], F.desc[i2])
            if d < best:
                best, bidx = d, i2
        if best <= max_dist:
            state[bidx] = i
            obs[bidx] = int(q.obs[i]) if q.obs is not None else 1
            nmatches += 1
            if check_ori:
                rot[_rot_bin(q.angle[i], F.angle[bidx], factor)].append(bidx)
    if check_ori:
        keep = three_maxima(rot)
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for idx in rot[b]:
                    state[idx] = -1
                    nmatches -= 1
    return nmatches, np.array(state, np.int32)


def stereo_unproject(kps_xy, depth, Tcw, fx, fy, cx, cy, bf):
    """Frame::ComputeStereoFromRGBD (src/Frame.cc:850-871) + Frame::UnprojectStereo (:879-899) for every keypoint:
    d = imDepth.at<float>(v, u) (float coordinates truncated by the int conversion), uRight = x - bf / d,
    x3Dc = ((u - cx) z invfx, (v - cy) z invfy, z) with invfx = 1.0f / fx (:215-216), world = mRwc x3Dc + mOw with
    mRwc = mRcw.t() and mOw = -mRwc mtcw (src/Frame.cc:373; cv::Mat products of evaluated matrices: float accumulation,
    '+ C' in double).  -> (uright, depth, xw, valid)."""
    fx, fy, cx, cy, bf = (f32(v) for v in (fx, fy, cx, cy, bf))
    invfx, invfy = f32(f32(1.0) / fx), f32(f32(1.0) / fy)
    T = np.asarray(Tcw, np.float32).reshape(4, 4)
    Rcw, tcw = T[:3, :3], T[:3, 3]
    # mOw = -mRwc*mtcw (src/Frame.cc:373): evaluated transpose -> small-matrix gemm, float accumulation
    Ow = [f32(-f32(f32(f32(Rcw[0, i] * tcw[0]) + f32(Rcw[1, i] * tcw[1])) + f32(Rcw[2, i] * tcw[2]))) for i in range(3)]
    n = len(kps_xy)
    ur = np.full(n, -1, np.float32)
    dp = np.full(n, -1, np.float32)
    xw = np.zeros((n, 3), np.float32)
    va = np.zeros(n, np.uint8)
    for i, (u, v) in enumerate(kps_xy):
        u, v = f32(u), f32(v)
        d = f32(depth[int(v), int(u)])
        if d > 0:
            dp[i] = d
            ur[i] = f32(u - f32(bf / d))
            x = f32(f32(f32(u - cx) * d) * invfx)
            y = f32(f32(f32(v - cy) * d) * invfy)
            for r in range(3):
                xw[i, r] = gemm3([Rcw[0, r], Rcw[1, r], Rcw[2, r]], [x, y, d], Ow[r])
            va[i] = 1
    return ur, dp, xw, va


def search_for_initialization(F1, F2, prev_xy, window=100, nnratio=0.9, check_ori=True):
    """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:523-660): level-0
    keypoints of F1 search F2 in a fixed window around their previous match; a strictly closer match steals an F2
    keypoint that is already matched; ratio test against the second best; orientation pruning only clears entries
    that are still matched.  Returns (nmatches, matches12, updated vbPrevMatched)."""
    nnratio = f32(nnratio)
    factor = f32(f32(1.0) / f32(HISTO_LENGTH))
    grid = Grid(F2)
    INT_MAX = 2147483647
    m12 = [-1] * F1.n
    m21 = [-1] * F2.n
    mdist = [INT_MAX] * F2.n
    prev = np.array(prev_xy, np.float32).reshape(-1, 2).copy()
    rot = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    for i1 in range(F1.n):
        level1 = int(F1.octave[i1])
        if level1 > 0:
            continue
        cand = grid.query(f32(prev[i1, 0]), f32(prev[i1, 1]), f32(window), level1, level1)
        if not cand:
            continue
        best, best2, bidx = INT_MAX, INT_MAX, -1
        for i2 in cand:
            d = hamming(F1.desc[i1], F2.desc[i2])
            if mdist[i2] <= d:
                continue
            if d < best:
                best2, best, bidx = best, d, i2
            elif d < best2:
                best2 = d
        if best <= TH_LOW and f32(best) < f32(f32(best2) * nnratio):
            if m21[bidx] >= 0:
                m12[m21[bidx]] = -1
                nmatches -= 1
            m12[i1], m21[bidx], mdist[bidx] = bidx, i1, best
            nmatches += 1
            if check_ori:
                rot[_rot_bin(F1.angle[i1], F2.angle[bidx], factor)].append(i1)
    if check_ori:
        keep = three_maxima(rot)
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for i1 in rot[b]:
                    if m12[i1] >= 0:
                        m12[i1] = -1
                        nmatches -= 1
    for i1 in range(F1.n):
        if m12[i1] >= 0:
            prev[i1] = (F2.x[m12[i1]], F2.y[m12[i1]])
    return nmatches, np.array(m12, np.int32), prev
