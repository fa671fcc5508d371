"""Reduced tracking loop for BASELINE.json configs[4] (System::TrackRGBD drop-in, ATE on identical inputs).

The reference's Tracking thread (src/Tracking.cc) needs g2o's pose optimiser, DBoW2 and the local map -- none of which is
on this path or in this image.  What CAN be exercised end to end is the path itself in the order Tracking drives it:
per frame  ORBextractor::operator()  ->  ComputeStereoFromRGBD / UnprojectStereo (map points of the last frame, in the
world frame of the LAST ESTIMATED pose)  ->  SearchByProjection(cur, last, th) with the motion-model prediction  ->  pose
update.  The pose update here is a plain Gauss-Newton on the reprojection error of the matches (numpy, float64,
deterministic; it stands in for Optimizer::PoseOptimization and is NOT a restatement of it), so that the loop closes and a
trajectory comes out; `backend` chooses who runs the path -- the B200 library or the CPU oracle -- and because the path is
bit-exact the two trajectories, and therefore their ATE, are identical."""
from __future__ import annotations

import numpy as np

from . import synth
from ._abi import FrameView, LastView


def _exp_se3(xi):
    """First-order-safe SE(3) exponential (rotation by Rodrigues, translation left-multiplied)."""
    w, v = xi[:3], xi[3:]
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        R, V = np.eye(3) + K, np.eye(3) + 0.5 * K
    else:
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * K @ K
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, V @ v
    return T


def solve_pose(T0, Xw, uv, fx, fy, cx, cy, iters=10, huber=np.sqrt(5.991)):
    """Gauss-Newton on the pixel reprojection error of world points Xw (n x 3) seen at uv (n x 2), from pose T0."""
    T = np.asarray(T0, np.float64).copy()
    Xw = np.asarray(Xw, np.float64)
    uv = np.asarray(uv, np.float64)
    for _ in range(iters):
        Xc = Xw @ T[:3, :3].T + T[:3, 3]
        z = Xc[:, 2]
        ok = z > 1e-3
        pu, pv = fx * Xc[:, 0] / z + cx, fy * Xc[:, 1] / z + cy
        r = np.stack([pu - uv[:, 0], pv - uv[:, 1]], 1)
        e = np.linalg.norm(r, axis=1)
        w = np.where(e <= huber, 1.0, huber / np.maximum(e, 1e-12)) * ok
        x, y, iz = Xc[:, 0], Xc[:, 1], 1.0 / np.where(ok, z, 1.0)
        J = np.zeros((len(Xw), 2, 6))
        J[:, 0, 0] = -fx * x * y * iz * iz; J[:, 0, 1] = fx * (1 + x * x * iz * iz); J[:, 0, 2] = -fx * y * iz
        J[:, 0, 3] = fx * iz; J[:, 0, 5] = -fx * x * iz * iz
        J[:, 1, 0] = -fy * (1 + y * y * iz * iz); J[:, 1, 1] = fy * x * y * iz * iz; J[:, 1, 2] = fy * x * iz
        J[:, 1, 4] = fy * iz; J[:, 1, 5] = -fy * y * iz * iz
        H = np.einsum("nij,n,nik->jk", J, w, J)
        b = np.einsum("nij,n,ni->j", J, w, r)
        dx = np.linalg.solve(H + 1e-9 * np.eye(6), -b)
        T = _exp_se3(dx) @ T
        if np.linalg.norm(dx) < 1e-10:
            break
    return T


class ReducedTracker:
    """extract / unproject / match are injected (B200 mirror classes or the oracle): see tests/test_tracking_*.py."""

    def __init__(self, extract, stereo_unproject, match, scale_factors, th=15.0):
        self.extract, self.unproject, self.match, self.sf, self.th = extract, stereo_unproject, match, scale_factors, th
        self.poses, self.nmatches = [], []
        self._last = None
        self._vel = np.eye(4)

    def track(self, gray, depth, T_init=None):
        fx, fy, cx, cy, bf = synth.FX, synth.FY, synth.CX, synth.CY, synth.BF
        K, D = self.extract(gray)
        if self._last is None:
            T = np.eye(4) if T_init is None else np.asarray(T_init, np.float64)
            self.nmatches.append(0)
        else:
            Kl, Dl, depth_l, Tl = self._last
            T_pred = (self._vel @ Tl).astype(np.float32)                              # motion model (src/Tracking.cc:1332)
            ur_c, _, _, _ = self.unproject(K, depth, T_pred)
            _, _, xw, valid = self.unproject(Kl, depth_l, Tl.astype(np.float32))
            cur = FrameView(K["x"], K["y"], K["octave"], K["angle"], ur_c, D, T_pred, fx, fy, cx, cy, bf, 0.0, float(gray.shape[1]),
                            0.0, float(gray.shape[0]), self.sf)
            last = LastView(xw, valid, Kl["octave"], Kl["angle"], Dl, Tl.astype(np.float32), mp_obs=np.ones(len(Kl), np.int32))
            n, c2l = self.match(cur, last, self.th)
            sel = np.nonzero(c2l >= 0)[0]
            self.nmatches.append(int(n))
            T = solve_pose(T_pred, xw[c2l[sel]], np.stack([K["x"][sel], K["y"][sel]], 1), fx, fy, cx, cy) if len(sel) >= 6 else T_pred.astype(np.float64)
            self._vel = T @ np.linalg.inv(Tl)
        self._last = (K, D, depth, np.asarray(T, np.float64))
        self.poses.append(np.asarray(T, np.float64))
        return self.poses[-1]

    def trajectory(self, stamps):
        """{stamp: [tx, ty, tz]} of the camera centres (Twc translation), the format evaluate_ate reads."""
        out = {}
        for s, T in zip(stamps, self.poses):
            R, t = T[:3, :3], T[:3, 3]
            out[float(s)] = (-R.T @ t).tolist()
        return out
