"""Synthetic, geometrically consistent inputs for the ORBmatcher members that work on KeyFrame / MapPoint graphs
(relocalisation and loop-closing projections, Fuse x2, SearchBySim3, SearchForTriangulation): world points in front of
two cameras, keypoints at their projections, descriptors that agree up to a few flipped bits."""
import numpy as np

from orb_slam2_ssd_semantic_b200 import synth
from orb_slam2_ssd_semantic_b200._abi import FrameView, QueriesView, TriKFView

SF = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2), np.float32)])).astype(np.float32)
FX, FY, CX, CY, BF = synth.FX, synth.FY, synth.CX, synth.CY, synth.BF


def pose(rng, trans=0.15, rot_deg=4.0):
    a = np.deg2rad(rng.normal(0, rot_deg, 3))
    cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = rng.normal(0, trans, 3)
    return T.astype(np.float32)


def project(T, X):
    Xc = X @ T[:3, :3].T.astype(np.float64) + T[:3, 3].astype(np.float64)
    return FX * Xc[:, 0] / Xc[:, 2] + CX, FY * Xc[:, 1] / Xc[:, 2] + CY, Xc[:, 2]


def world_points(rng, n):
    z = rng.uniform(1.0, 5.0, n)
    return np.stack([(rng.uniform(20, 620, n) - CX) * z / FX, (rng.uniform(20, 460, n) - CY) * z / FY, z], 1)


def noisy(rng, desc, nbytes=3, frac=1.0):
    d = desc.copy()
    sel = rng.random(len(d)) < frac
    d[sel, :nbytes] ^= rng.integers(0, 256, size=(int(sel.sum()), nbytes), dtype=np.uint8)
    return d


def frame_of(rng, T, X, desc, stereo_frac=0.7, jitter=0.6, octaves=None, extra=60):
    """Keypoints at the projections of X (+ `extra` unrelated ones); returns (FrameView, index of each X or -1)."""
    u, v, z = project(T, X)
    ok = (z > 0.2) & (u > 3) & (u < 637) & (v > 3) & (v < 477)
    idx = np.nonzero(ok)[0]
    n = len(idx) + extra
    x = np.concatenate([u[idx] + rng.normal(0, jitter, len(idx)), rng.uniform(5, 635, extra)]).astype(np.float32)
    y = np.concatenate([v[idx] + rng.normal(0, jitter, len(idx)), rng.uniform(5, 475, extra)]).astype(np.float32)
    zz = np.concatenate([z[idx], rng.uniform(1, 5, extra)])
    octv = (rng.integers(0, 4, n) if octaves is None else np.concatenate([octaves[idx], rng.integers(0, 4, extra)])).astype(np.int32)
    d = np.concatenate([noisy(rng, desc[idx]), rng.integers(0, 256, size=(extra, 32), dtype=np.uint8)])
    ur = np.where(rng.random(n) < stereo_frac, x - BF / zz, -1).astype(np.float32)
    perm = rng.permutation(n)
    # the same world point keeps (nearly) the same orientation in every frame, so the rotation histogram has a peak
    ang = np.concatenate([(idx * 37.0) % 360.0 + rng.normal(0, 3.0, len(idx)), rng.uniform(0, 360, extra)]) % 360.0
    F = FrameView(x[perm], y[perm], octv[perm], ang.astype(np.float32)[perm], ur[perm], d[perm], T, FX, FY, CX,
                  CY, BF, 0, 640, 0, 480, SF)
    owner = np.concatenate([idx, np.full(extra, -1)])[perm]
    return F, owner


def map_points(rng, X, desc, Ow, oracle, bad_frac=0.03, null_frac=0.05, obs_max=4, octaves=None):
    """octaves: pyramid level the point was seen at -> mfMaxDistance = dist * 1.2^octave (MapPoint::UpdateNormalAndDepth),
    so that PredictScale lands on the keypoints' levels; None = unrelated distance bounds."""
    n = len(X)
    PO = X - Ow
    dist = np.linalg.norm(PO, axis=1)
    if octaves is not None:
        maxd = dist * SF[octaves].astype(np.float64) * 1.03
        normal = (PO / dist[:, None] + rng.normal(0, 0.05, (n, 3)))
        normal /= np.linalg.norm(normal, axis=1)[:, None]
        return oracle.MapPointsView(X.astype(np.float32), noisy(rng, desc, 2, 0.7), valid=(rng.random(n) >= null_frac).astype(np.uint8),
                                    bad=(rng.random(n) < bad_frac).astype(np.uint8), normal=normal.astype(np.float32),
                                    min_dist=(maxd / SF[7]).astype(np.float32), max_dist=maxd.astype(np.float32),
                                    obs=rng.integers(0, obs_max, n).astype(np.int32))
    normal = (PO / dist[:, None] + rng.normal(0, 0.05, (n, 3)))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    return oracle.MapPointsView(X.astype(np.float32), noisy(rng, desc, 2, 0.7), valid=(rng.random(n) >= null_frac).astype(np.uint8),
                                bad=(rng.random(n) < bad_frac).astype(np.uint8), normal=normal.astype(np.float32),
                                min_dist=(dist * rng.uniform(0.3, 0.9, n)).astype(np.float32),
                                max_dist=(dist * rng.uniform(1.1, 3.0, n)).astype(np.float32),
                                obs=rng.integers(0, obs_max, n).astype(np.int32))


def camera_centre(T):
    return -(T[:3, :3].astype(np.float64).T @ T[:3, 3].astype(np.float64))


def best_queries(rng, F, m=250):
    """Random per-query windows on a frame for the BEST search."""
    n = F.n
    sel = rng.integers(0, n, m)
    lvl = np.clip(F.octave[sel] + rng.integers(-1, 2, m), 0, 7)
    d2 = F.desc[sel].copy()
    d2[:, :3] ^= rng.integers(0, 256, size=(m, 3), dtype=np.uint8)
    u = (F.x[sel] + rng.normal(0, 4, m)).astype(np.float32)
    v = (F.y[sel] + rng.normal(0, 4, m)).astype(np.float32)
    return QueriesView((rng.random(m) < 0.9).astype(np.uint8), u, v, (np.float32(6.0) * SF[lvl]).astype(np.float32), lvl - 1, lvl,
                       d2, np.zeros(m, np.float32), uright=(u - BF / rng.uniform(1, 5, m)).astype(np.float32))


def fundamental(T1, T2):
    """F12 with x1^T F12 x2 = 0, as LocalMapping::ComputeF12 builds it (K^-T [t12]x R12 K^-1)."""
    R1, t1, R2, t2 = T1[:3, :3].astype(np.float64), T1[:3, 3].astype(np.float64), T2[:3, :3].astype(np.float64), T2[:3, 3].astype(np.float64)
    R12 = R1 @ R2.T
    t12 = -R1 @ R2.T @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]])
    Ki = np.linalg.inv(K)
    return (Ki.T @ tx @ R12 @ Ki).astype(np.float32)


def tri_pair(rng, n=400, mono_frac=0.6):
    """Two keyframes observing the same points, BoW nodes shared by corresponding keypoints."""
    X = world_points(rng, n)
    desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    T1, T2 = pose(rng, 0.05, 1.0), pose(rng, 0.25, 3.0)
    F1, o1 = frame_of(rng, T1, X, desc, stereo_frac=1 - mono_frac, jitter=0.4)
    F2, o2 = frame_of(rng, T2, X, desc, stereo_frac=1 - mono_frac, jitter=0.4)
    nodes = rng.integers(0, 25, n) * 7

    def fv(owner):
        d = {}
        for i, o in enumerate(owner):
            node = int(nodes[o]) if o >= 0 else int(rng.integers(0, 25)) * 7
            if rng.random() < 0.05:
                node += 1          # a few keypoints land in a node the other keyframe does not have
            d.setdefault(node, []).append(i)
        return d
    k1 = TriKFView(F1.x, F1.y, F1.octave, F1.angle, F1.uright, F1.desc, (rng.random(F1.n) < 0.3).astype(np.uint8), fv(o1))
    k2 = TriKFView(F2.x, F2.y, F2.octave, F2.angle, F2.uright, F2.desc, (rng.random(F2.n) < 0.3).astype(np.uint8), fv(o2))
    return k1, k2, T1, T2, F1, fundamental(T1, T2)


def scene(rng, oracle, n=500):
    """Two keyframes (KF, KF2) + a current frame observing n world points, the keyframes' MapPoints, a list of candidate
    MapPoints `pts` (the same world points, slightly moved), a Sim3 pose S of KF and the relative pose of the keyframes."""
    X = world_points(rng, n)
    desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    T = pose(rng)
    octs = rng.integers(0, 4, n)
    KF, owner = frame_of(rng, T, X, desc, octaves=octs)
    has = (owner >= 0) & (rng.random(KF.n) < 0.6)
    kf_mps = map_points(rng, np.where(has[:, None], X[np.maximum(owner, 0)], 0.0), KF.desc, camera_centre(T), oracle, octaves=KF.octave)
    kf_mps.valid = has.astype(np.uint8)
    pts = map_points(rng, X + rng.normal(0, 0.004, X.shape), desc, camera_centre(T), oracle, octaves=octs)
    S = T.copy()
    S[:3, :] *= np.float32(1.3)
    cur, _ = frame_of(rng, pose(rng), X, desc, octaves=octs)
    cur.mp_obs = np.where(rng.random(cur.n) < 0.1, 1, -1).astype(np.int32)
    T2 = pose(rng, 0.2, 3.0)
    KF2, owner2 = frame_of(rng, T2, X, desc, octaves=octs)
    has2 = (owner2 >= 0) & (rng.random(KF2.n) < 0.6)
    mp2 = map_points(rng, np.where(has2[:, None], X[np.maximum(owner2, 0)], 0.0), KF2.desc, camera_centre(T2), oracle, octaves=KF2.octave)
    mp2.valid = has2.astype(np.uint8)
    T12 = T.astype(np.float64) @ np.linalg.inv(T2.astype(np.float64))
    return dict(KF=KF, kf_mps=kf_mps, pts=pts, S=S, cur=cur, KF2=KF2, mp2=mp2, R12=T12[:3, :3], t12=T12[:3, 3],
                in_kf=(rng.random(n) < 0.05).astype(np.uint8), found=(rng.random(KF.n) < 0.1).astype(np.uint8),
                matched=np.where(rng.random(KF.n) < 0.1, -2, -1).astype(np.int32))
