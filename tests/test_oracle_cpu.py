"""CPU suite: pins the C++ oracle (test infrastructure) against (a) the committed golden vectors generated with the
real OpenCV primitives / the pure-Python restatements, (b) cv2 itself when it is importable, (c) artefacts of the
reference (map.bin keyframes, octomap.ot log-odds lattice)."""
import glob
import os

import numpy as np
import pytest

from orb_slam2_ssd_semantic_b200 import synth
from orb_slam2_ssd_semantic_b200._abi import FrameView, LastView

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _views(z):
    fx, fy, cx, cy, bf = [float(v) for v in z["cam"]]
    cur = FrameView(z["cur_x"], z["cur_y"], z["cur_oct"], z["cur_angle"], z["cur_uright"], z["cur_desc"], z["cur_Tcw"], fx,
                    fy, cx, cy, bf, 0.0, 640.0, 0.0, 480.0, z["sf"])
    last = LastView(z["last_xw"], z["last_valid"], z["last_oct"], z["last_angle"], z["last_desc"], z["last_Tcw"],
                    mp_obs=np.ones(len(z["last_valid"]), np.int32))
    return cur, last


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "extract_*.npz"))))
def test_extractor_oracle_reproduces_cv2_golden(oracle, path):
    z = np.load(path)
    prm = z["params"]
    R = oracle.RefExtractor(int(prm[0]), float(prm[1]), int(prm[2]), int(prm[3]), int(prm[4]))
    K, D = R(z["image"])
    assert (R.candidates_per_level == z["candidates"]).all()
    assert K.tobytes() == z["kps"].tobytes()
    assert (D == z["desc"]).all()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "match_mapbin_*.npz"))))
def test_matcher_oracle_reproduces_mapbin_golden(oracle, path):
    z = np.load(path)
    cur, last = _views(z)
    n, m = oracle.search_by_projection_last(cur, last, float(z["th"]), False, 0.9, True)
    assert n == int(z["nmatches"]) and n > 100
    assert (m == z["cur2last"]).all()


def test_matcher_oracle_vs_python_restatement_random(oracle):
    """Small random frames: C++ oracle == pure-Python restatement, incl. pre-existing points, obs=0 double claims,
    no-orientation and mono variants."""
    from oracle import match_py
    rng = np.random.default_rng(5)
    sf = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2), np.float32)])).astype(np.float32)
    for case in range(12):
        n = int(rng.integers(40, 160))
        x = rng.uniform(5, 635, n).astype(np.float32)
        y = rng.uniform(5, 475, n).astype(np.float32)
        octv = rng.integers(0, 8, n).astype(np.int32)
        desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        T = np.eye(4, dtype=np.float32)
        T[:3, 3] = rng.normal(0, 0.02, 3)
        if case % 4 == 1:
            T[2, 3] = 0.3
        if case % 4 == 2:
            T[2, 3] = -0.3
        z = rng.uniform(0.5, 4.0, n).astype(np.float32)
        ur = np.where(rng.random(n) < 0.7, x - synth.BF / z, -1).astype(np.float32)
        cur = FrameView(x, y, octv, rng.uniform(0, 360, n).astype(np.float32), ur, desc, T, synth.FX, synth.FY, synth.CX,
                        synth.CY, synth.BF, 0, 640, 0, 480, sf)
        if case % 3 == 0:
            cur.mp_obs = rng.integers(-1, 2, n).astype(np.int32)
        m = int(rng.integers(40, 160))
        sel = rng.integers(0, n, m)
        zz = z[sel] * rng.uniform(0.98, 1.02, m).astype(np.float32)
        xw = np.stack([(x[sel] + rng.normal(0, 3, m) - synth.CX) * zz / synth.FX,
                       (y[sel] + rng.normal(0, 3, m) - synth.CY) * zz / synth.FY, zz], 1).astype(np.float32)
        d2 = desc[sel].copy()
        d2[:, :4] ^= rng.integers(0, 256, size=(m, 4), dtype=np.uint8)
        last = LastView(xw, (rng.random(m) < 0.9).astype(np.uint8), np.clip(octv[sel] + rng.integers(-1, 2, m), 0, 7),
                        rng.uniform(0, 360, m).astype(np.float32), d2, np.eye(4, dtype=np.float32),
                        mp_obs=rng.integers(0, 2, m).astype(np.int32))
        chk = case % 5 != 4
        mono = case == 7
        n1, m1 = oracle.search_by_projection_last(cur, last, 15.0, mono, 0.9, chk)
        n2, m2 = match_py.search_by_projection_last(cur, last, 15.0, mono, chk)
        assert n1 == n2 and (m1 == m2).all(), "case %d" % case


def test_primitives_against_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (h, w) in [(480, 640), (97, 131), (33, 47)]:
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        dw, dh = int(round(w / 1.2)), int(round(h / 1.2))
        assert (cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR) == oracle.resize(img, dw, dh)).all()
        blur = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert (blur == oracle.blur(img)).all()
    img = synth.synth_frame(3, 1)
    for t in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True)
        for (y0, x0, hh, ww) in [(16, 16, 38, 37), (100, 200, 37, 36), (0, 0, 60, 60), (400, 500, 9, 9)]:
            roi = img[y0:y0 + hh, x0:x0 + ww]
            k = det.detect(roi)
            f = oracle.fast(roi, t)
            assert len(k) == len(f)
            assert all((int(a.pt[0]), int(a.pt[1]), int(a.response)) == tuple(b) for a, b in zip(k, f))
    for _ in range(2000):
        yv, xv = [float(v) for v in rng.integers(-200000, 200000, 2)]
        assert oracle.fast_atan2(yv, xv) == np.float32(cv2.fastAtan2(yv, xv))


def test_full_extractor_against_cv2_restatement(oracle):
    pytest.importorskip("cv2")
    from oracle.orb_cv2 import ORBextractorCV2
    img = synth.synth_frame(21, 4, h=240, w=320)
    K, D = oracle.RefExtractor(400, 1.2, 6, 20, 7)(img)
    K2, D2 = ORBextractorCV2(400, 1.2, 6, 20, 7)(img)
    assert K.tobytes() == K2.tobytes() and (D == D2).all()


def test_scale_tables(oracle):
    R = oracle.RefExtractor(1000, 1.2, 8, 20, 7)
    assert R.mnFeaturesPerLevel.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert oracle.RefExtractor(2000, 1.2, 8, 20, 7).mnFeaturesPerLevel.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert R.umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    R(synth.synth_frame(1, 0))
    assert [R.level(l).shape for l in range(8)] == [(480, 640), (400, 533), (333, 444), (278, 370), (231, 309),
                                                    (193, 257), (161, 214), (134, 179)]


def test_octomap_artifact_pins_logodds_lattice(oracle):
    """Every node value stored in the reference's octomap.ot is reachable by the oracle's clamp-add update rule with
    its (hit, miss, clamp) constants -> pins setProbHit/Miss/ClampingThres and float32 storage."""
    z = np.load(os.path.join(G, "octomap_logodds.npz"))
    hit, miss, cmin, cmax = oracle.RefOccupancy().constants()
    assert abs(hit - 0.8473) < 1e-4 and abs(miss + 0.405465) < 1e-6
    assert abs(cmin + 1.99243) < 1e-5 and abs(cmax - 3.4761) < 1e-4
    reach = {np.float32(0.0)}
    frontier = [np.float32(0.0)]
    while frontier and len(reach) < 20000:
        nxt = []
        for v in frontier:
            for d in (hit, miss):
                w = np.float32(min(max(np.float32(v + d), cmin), cmax))
                if w not in reach:
                    reach.add(w)
                    nxt.append(w)
        frontier = nxt
    r = np.array(sorted(reach), np.float32)
    vals = z["values"]
    assert vals.min() >= cmin - 1e-6 and vals.max() <= cmax + 1e-6
    d = np.abs(vals[:, None] - r[None, :]).min(axis=1)
    assert d.max() < 2e-6, "octomap.ot holds values off the oracle's log-odds lattice: %s" % vals[d.argmax()]


def test_occupancy_oracle_properties(oracle):
    ws = synth.WallStream(seed=5, n=2)
    gray, depth, rgb, T = ws.frame(0)
    R = oracle.RefOccupancy()
    n = R.insert_keyframe(T, depth, rgb, synth.FX, synth.FY, synth.CX, synth.CY)
    pts, col, lab = R.last_points()
    assert n == len(pts) > 1000
    # the wall is the plane z_w = 2: every centroid lies on it; leaf filter keeps < 1 point per cm^3 cell
    assert np.abs(pts[:, 2] - 2.0).max() < 2e-3
    keys, lo = R.export_leaves()
    assert len(np.unique(keys, axis=0)) == len(keys) and np.allclose(lo, R.constants()[0])
    # ray: end cell excluded, first cell is the origin's, consecutive cells are face neighbours
    ray = R.ray([0.01, 0.02, 0.03], [1.234, -0.5, 2.2]).astype(np.int64)
    assert (ray[0] == 32768).all()
    assert (np.abs(np.diff(ray, axis=0)).sum(axis=1) == 1).all()
    end = (np.floor(np.array([1.234, -0.5, 2.2]) / 0.05) + 32768).astype(np.int64)
    assert not (ray == end).all(axis=1).any() and np.abs(ray[-1] - end).sum() == 1


def test_restated_glibc_sincosf_equals_host_libm(oracle):
    """The CUDA kernel runs include/glibc_sincosf.h; the oracle calls the host libm like the reference does
    (src/ORBextractor.cc:97).  They must agree on every float angle the descriptor path can see ([0, 2 pi])."""
    import ctypes as C
    L = oracle.lib()
    L.orb_ref_sincosf_mismatches.restype = C.c_long
    L.orb_ref_sincosf_mismatches.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    two_pi_bits = int(np.float32(6.2831855).view(np.uint32))
    assert L.orb_ref_sincosf_mismatches(0, two_pi_bits + 64, 53) == 0        # ~20 M angles, every binade
    assert L.orb_ref_sincosf_mismatches(int(np.float32(3.0).view(np.uint32)), two_pi_bits, 1) == 0   # dense top binade


def test_ot_export_round_trips_reference_octomap():
    """octree_io.build_ot(leaves of the reference's octomap.ot) reproduces that file: same node count and pre-order,
    every log-odds value, every child mask, every leaf colour (inner colours are history dependent in octomap and
    are written white -- see octree_io.py)."""
    from orb_slam2_ssd_semantic_b200 import octree_io as O
    z = np.load(os.path.join(G, "octomap_nodes.npz"))
    nodes = np.zeros(len(z["v"]), O.NODE_DT)
    nodes["v"], nodes["rgb"], nodes["child"] = z["v"], z["rgb"], z["child"]
    keys, depths, vals, cols, consistent, used = O.leaves_from_nodes(nodes)
    assert consistent and used == len(nodes)               # inner value == max over children everywhere
    assert np.bincount(depths, minlength=17)[13:].tolist() == [4, 187, 8376, 310363]   # pruned leaves exist
    k16, v16, c16 = O.expand_to_max_depth(keys, depths, vals, cols)
    out = O.build_ot(k16, v16, c16, 0.05, "0.05")
    res, n2, hdr = O.read_ot(out)
    assert hdr == str(z["header"]) and len(n2) == len(nodes) == 390133
    assert (n2["v"] == nodes["v"]).all() and (n2["child"] == nodes["child"]).all()
    leaf = nodes["child"] == 0
    assert (n2["rgb"][leaf] == nodes["rgb"][leaf]).all()


def test_global_refilter_against_numpy_restatement():
    """occ_ref_global_refilter (PCL VoxelGrid over the accumulated T-variant map) against an independent numpy
    restatement of the same published algorithm: bounding box -> min_b/div_b -> linear index -> stable sort -> float
    sums in order."""
    from oracle import ref
    rng = np.random.default_rng(3)
    xyz = rng.uniform(-1.5, 2.5, (6000, 3)).astype(np.float32)
    xyz[17] = np.nan
    xyz[400, 1] = np.inf
    rgb = rng.integers(0, 256, (6000, 3)).astype(np.uint8)
    leaf = np.float32(0.13)
    o, c = ref.global_refilter(xyz, rgb, leaf)
    inv = np.float32(1.0) / leaf
    fin = np.isfinite(xyz).all(1)
    P, Cc = xyz[fin], rgb[fin]
    mb = np.floor(P.min(0) * inv).astype(np.int64)
    div = np.floor(P.max(0) * inv).astype(np.int64) - mb + 1
    ijk = (np.floor(P * inv) - mb.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    u, st, cnt = np.unique(idx[order], return_index=True, return_counts=True)
    assert len(u) == len(o) and len(o) > 500
    for k in range(0, len(u), 37):
        acc = np.zeros(3, np.float32)
        accc = np.zeros(3, np.float32)
        for j in order[st[k]:st[k] + cnt[k]]:
            acc = acc + P[j]
            accc = accc + Cc[j].astype(np.float32)
        assert (o[k] == acc / np.float32(cnt[k])).all()
        assert (c[k] == (accc / np.float32(cnt[k])).astype(np.uint8)).all()


def test_occupancy_oracle_against_python_restatement(oracle):
    """oracle/occ_ref.cpp against oracle/occ_py.py (second restatement, written from the reference text and SURVEY
    App. A.6/A.7): points, leaf keys and clamped log-odds after three keyframes of a small image, with ground-labelled
    pixels casting rays."""
    from oracle.occ_py import PyOccupancy
    rng = np.random.default_rng(11)
    rows, cols = 40, 56
    fx, fy, cx, cy = 48.0, 47.0, 27.6, 19.3
    ref = oracle.RefOccupancy()
    py = PyOccupancy()
    for k in range(3):
        yy, xx = np.mgrid[0:rows, 0:cols]
        depth = (1.4 + 0.5 * np.sin(xx / 9.0 + k) * np.cos(yy / 7.0) + 0.02 * rng.standard_normal((rows, cols))).astype(np.float32)
        depth[3, 5] = 0.2          # below depth_min
        depth[10:12, 20:30] = 3.4  # above depth_max
        rgb = rng.integers(0, 256, (rows, cols, 3)).astype(np.uint8)
        label = np.zeros((rows, cols), np.uint8)
        label[rows // 2:, :] = 1   # lower half is "ground": rays carve free space
        a = 0.05 * k
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
        T[:3, 3] = np.array([0.03 * k, -0.01 * k, 0.02 * k], np.float32)
        n_ref = ref.insert_keyframe(T, depth, rgb, fx, fy, cx, cy, label)
        n_py = py.insert_keyframe(T, depth, fx, fy, cx, cy, label)
        pts, _, lab = ref.last_points()
        assert n_ref == n_py == len(pts) > 500
        assert (pts == py.points).all() and (lab == py.labels).all()
    keys, lo = ref.export_leaves()
    got = {tuple(int(v) for v in k): np.float32(x) for k, x in zip(keys, lo)}
    assert len(got) == len(py.leaves) > 300
    assert all(got[k] == py.leaves[k] for k in got)
    assert sum(1 for v in got.values() if v < 0) > 50       # carved free cells exist


def test_other_searches_oracle_vs_python_restatements(oracle):
    """oracle/match_ref.cpp against oracle/match_py.py (second restatement from the reference text) for
    SearchByProjection(Frame&, vector<MapPoint*>&, th) (:63-156), SearchByBoW(KF,F) (:217-363) and SearchByBoW(KF,KF)
    (:665-812) on random frames: pre-existing points, stereo gate, ratio test, orientation pruning on/off."""
    from oracle import match_py
    from orb_slam2_ssd_semantic_b200._abi import BowView, TrackPointsView
    rng = np.random.default_rng(21)
    sf = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2), np.float32)])).astype(np.float32)
    tot = [0, 0, 0]
    for case in range(6):
        n = int(rng.integers(80, 200))
        x = rng.uniform(5, 635, n).astype(np.float32)
        y = rng.uniform(5, 475, n).astype(np.float32)
        octv = rng.integers(0, 8, n).astype(np.int32)
        desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        z = rng.uniform(0.5, 4, n).astype(np.float32)
        ur = np.where(rng.random(n) < 0.7, x - synth.BF / z, -1).astype(np.float32)
        F = FrameView(x, y, octv, rng.uniform(0, 360, n).astype(np.float32), ur, desc, np.eye(4, dtype=np.float32), synth.FX,
                      synth.FY, synth.CX, synth.CY, synth.BF, 0, 640, 0, 480, sf)
        if case % 2 == 0:
            F.mp_obs = rng.integers(-1, 2, n).astype(np.int32)
        m = int(rng.integers(80, 200))
        sel = rng.integers(0, n, m)
        d2 = desc[sel].copy()
        d2[:, :3] ^= rng.integers(0, 256, size=(m, 3), dtype=np.uint8)
        px = (x[sel] + rng.normal(0, 4, m)).astype(np.float32)
        py = (y[sel] + rng.normal(0, 4, m)).astype(np.float32)
        pts = TrackPointsView((rng.random(m) < 0.9).astype(np.uint8), px, py, (px - synth.BF / z[sel]).astype(np.float32),
                              np.clip(octv[sel] + rng.integers(0, 2, m), 0, 7), rng.uniform(0.99, 1.0, m).astype(np.float32), d2,
                              mp_obs=rng.integers(0, 2, m).astype(np.int32))
        for th in (1.0, 3.0):
            a = oracle.search_by_projection_points(F, pts, th, 0.8)
            b = match_py.search_by_projection_points(F, pts, th, 0.8)
            assert a[0] == b[0] and (a[1] == b[1]).all(), ("points", case, th)
            tot[0] += a[0]
        nw = int(rng.integers(4, 40))
        fv1, fv2 = {}, {}
        for i, w in enumerate(rng.integers(0, nw, n)):
            fv1.setdefault(int(w) * 3, []).append(i)
        for i, w in enumerate(rng.integers(0, nw, m)):
            fv2.setdefault(int(w) * 3 + (0 if rng.random() < 0.8 else 1), []).append(i)
        ang2 = rng.uniform(0, 360, m).astype(np.float32)
        K = BowView(desc, F.angle, fv1, valid=(rng.random(n) < 0.85).astype(np.uint8))
        Fr = BowView(d2, ang2, fv2)
        K2 = BowView(d2, ang2, fv2, valid=(rng.random(m) < 0.85).astype(np.uint8))
        for ori in (True, False):
            a, b = oracle.search_by_bow(K, Fr, 0.7, ori), match_py.search_by_bow(K, Fr, 0.7, ori)
            assert a[0] == b[0] and (a[1] == b[1]).all(), ("bow", case, ori)
            tot[1] += a[0]
            a, b = oracle.search_by_bow_kf(K, K2, 0.75, ori), match_py.search_by_bow_kf(K, K2, 0.75, ori)
            assert a[0] == b[0] and (a[1] == b[1]).all(), ("bow_kf", case, ori)
            tot[2] += a[0]
    assert min(tot) > 20, tot


def test_generic_projected_search_oracle_vs_python_restatement(oracle):
    """match_ref_projected (the loop body the relocalisation / loop-closing projection overloads share) against
    oracle/match_py.search_projected: both claim rules, with and without the stereo gate / orientation pruning."""
    from oracle import match_py
    from orb_slam2_ssd_semantic_b200._abi import QueriesView
    rng = np.random.default_rng(33)
    sf = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2), np.float32)])).astype(np.float32)
    tot = 0
    for case in range(8):
        n = int(rng.integers(80, 220))
        x = rng.uniform(5, 635, n).astype(np.float32)
        y = rng.uniform(5, 475, n).astype(np.float32)
        octv = rng.integers(0, 8, n).astype(np.int32)
        desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        z = rng.uniform(0.5, 4, n).astype(np.float32)
        ur = np.where(rng.random(n) < 0.7, x - synth.BF / z, -1).astype(np.float32)
        F = FrameView(x, y, octv, rng.uniform(0, 360, n).astype(np.float32), ur, desc, np.eye(4, dtype=np.float32), synth.FX,
                      synth.FY, synth.CX, synth.CY, synth.BF, 0, 640, 0, 480, sf)
        if case % 2 == 0:
            F.mp_obs = rng.integers(-1, 2, n).astype(np.int32)
        m = int(rng.integers(80, 220))
        sel = rng.integers(0, n, m)
        d2 = desc[sel].copy()
        d2[:, :3] ^= rng.integers(0, 256, size=(m, 3), dtype=np.uint8)
        u = (x[sel] + rng.normal(0, 4, m)).astype(np.float32)
        v = (y[sel] + rng.normal(0, 4, m)).astype(np.float32)
        lvl = np.clip(octv[sel] + rng.integers(-1, 2, m), 0, 7)
        for rule, (md, ori) in enumerate([(100, True), (64, False)]):
            q = QueriesView((rng.random(m) < 0.9).astype(np.uint8), u, v, (np.float32(8.0) * sf[lvl]).astype(np.float32),
                            lvl - 1, lvl + 1, d2, rng.uniform(0, 360, m).astype(np.float32),
                            uright=(u - synth.BF / z[sel]).astype(np.float32) if case % 3 == 0 else None,
                            obs=rng.integers(0, 2, m).astype(np.int32) if rule == 0 else None)
            a = oracle.search_projected(F, q, md, rule, ori)
            b = match_py.search_projected(F, q, md, rule, ori)
            assert a[0] == b[0] and (a[1] == b[1]).all(), (case, rule)
            tot += a[0]
    assert tot > 300


def test_stereo_unproject_oracle_vs_python_restatement(oracle):
    """frame_ref_stereo_unproject (ComputeStereoFromRGBD + UnprojectStereo, src/Frame.cc:850-899) against the second
    restatement in oracle/match_py.py, bit for bit, including keypoints without depth."""
    from oracle import match_py
    ws = synth.WallStream(seed=4, n=2)
    rng = np.random.default_rng(2)
    for t in range(2):
        gray, depth, rgb, T = ws.frame(t * 9)
        depth = depth.copy()
        depth[rng.random(depth.shape) < 0.3] = 0
        K, _ = oracle.RefExtractor(500, 1.2, 8, 20, 7)(gray)
        a = oracle.stereo_unproject(K, depth, T, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF)
        b = match_py.stereo_unproject(list(zip(K["x"], K["y"])), depth, T, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF)
        assert 0.5 * len(K) < a[3].sum() < 0.9 * len(K)
        for x, y in zip(a, b):
            assert np.asarray(x).tobytes() == np.asarray(y).tobytes()


def test_search_for_initialization_oracle_vs_python_restatement(oracle):
    """match_ref_initialization (SearchForInitialization, src/ORBmatcher.cc:523-660; oracle only so far, the kernel is
    next round's work) against the second restatement: steal-if-closer rule, ratio test, prune of still-matched entries."""
    from oracle import match_py
    rng = np.random.default_rng(41)
    sf = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2), np.float32)])).astype(np.float32)
    tot = 0
    for case in range(8):
        n = int(rng.integers(150, 400))
        x = rng.uniform(5, 635, n).astype(np.float32)
        y = rng.uniform(5, 475, n).astype(np.float32)
        octv = rng.integers(0, 3, n).astype(np.int32)           # a third of the keypoints on level 0
        desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        ang = rng.uniform(0, 360, n).astype(np.float32)
        mk = lambda xx, yy, dd, aa: FrameView(xx, yy, octv, aa, np.full(n, -1, np.float32), dd, np.eye(4, dtype=np.float32),
                                              synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, 0, 640, 0, 480, sf)
        F1 = mk(x, y, desc, ang)
        d2 = desc.copy()
        flip = rng.integers(0, 256, size=(n, 2), dtype=np.uint8)
        d2[:, :2] ^= flip
        dup = rng.integers(0, n, n // 6)                        # near-duplicate descriptors: contested F2 keypoints
        d2[dup] = d2[(dup + 1) % n]
        F2 = mk((x + rng.normal(0, 6, n)).astype(np.float32), (y + rng.normal(0, 6, n)).astype(np.float32), d2,
                (ang + rng.normal(0, 4, n)).astype(np.float32) % np.float32(360))
        prev = np.stack([x, y], 1)
        for window, ori in ((100, True), (30, False)):
            a = oracle.search_for_initialization(F1, F2, prev, window, 0.9, ori)
            b = match_py.search_for_initialization(F1, F2, prev, window, 0.9, ori)
            assert a[0] == b[0] and (a[1] == b[1]).all() and a[2].tobytes() == b[2].tobytes(), (case, window, ori)
            assert a[0] == (a[1] >= 0).sum()
            tot += a[0]
    assert tot > 200
