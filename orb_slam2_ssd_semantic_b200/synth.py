"""Deterministic synthetic inputs: 640x480 gray frames, RGB-D keyframes, camera streams.

There is no dataset in the sandbox, so every test and the bench draw their inputs from here.  Pure numpy
(PCG64), bit-reproducible across machines for a given numpy version; golden fixtures store the generated
arrays themselves so they do not depend on that.
"""
from __future__ import annotations

import numpy as np

W, H = 640, 480
# perfect/Examples/RGB-D/TUM3.yaml:8-25
FX, FY, CX, CY, BF = 535.4, 539.2, 320.1, 247.6, 40.0
DEPTH_FACTOR = 5000.0


def _bilinear_up(small: np.ndarray, h: int, w: int) -> np.ndarray:
    sh, sw = small.shape
    ys = (np.arange(h) + 0.5) * sh / h - 0.5
    xs = (np.arange(w) + 0.5) * sw / w - 0.5
    y0 = np.clip(np.floor(ys).astype(int), 0, sh - 1)
    x0 = np.clip(np.floor(xs).astype(int), 0, sw - 1)
    y1 = np.clip(y0 + 1, 0, sh - 1)
    x1 = np.clip(x0 + 1, 0, sw - 1)
    fy = np.clip(ys - y0, 0, 1)[:, None]
    fx = np.clip(xs - x0, 0, 1)[None, :]
    s = small.astype(np.float64)
    top = s[y0][:, x0] * (1 - fx) + s[y0][:, x1] * fx
    bot = s[y1][:, x0] * (1 - fx) + s[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


def texture(seed: int = 1234, h: int = H, w: int = W, nrect: int = 400, nline: int = 200) -> np.ndarray:
    """Static scene texture (float64, 0..255): smooth noise + rectangles + line segments."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = _bilinear_up(rng.integers(0, 256, size=(max(h // 16, 2), max(w // 16, 2))).astype(np.uint8), h, w)
    rng2 = np.random.Generator(np.random.PCG64(seed + 1))
    for _ in range(nrect):
        rw, rh = rng2.integers(8, 65, size=2)
        x = int(rng2.integers(0, max(w - 8, 1)))
        y = int(rng2.integers(0, max(h - 8, 1)))
        base[y:y + rh, x:x + rw] = float(rng2.integers(0, 256))
    for _ in range(nline):
        x0, x1 = rng2.integers(0, w, size=2)
        y0, y1 = rng2.integers(0, h, size=2)
        n = int(max(abs(int(x1) - int(x0)), abs(int(y1) - int(y0)), 1))
        xs = np.linspace(x0, x1, n + 1).round().astype(int)
        ys = np.linspace(y0, y1, n + 1).round().astype(int)
        base[ys, xs] = float(rng2.integers(0, 256))
    return base


_TEX_CACHE: dict = {}


def synth_frame(seed: int = 1234, t: int = 0, h: int = H, w: int = W) -> np.ndarray:
    """640x480 u8 frame: static texture + per-frame N(0, 2^2) noise (seed 10^6 + t)."""
    key = (seed, h, w)
    if key not in _TEX_CACHE:
        _TEX_CACHE[key] = texture(seed, h, w)
    rng = np.random.Generator(np.random.PCG64(10 ** 6 + t))
    img = _TEX_CACHE[key] + rng.normal(0.0, 2.0, size=(h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def adversarial_frames(h: int = H, w: int = W) -> dict:
    """Edge-case images the parity tests run."""
    rng = np.random.Generator(np.random.PCG64(99))
    yy, xx = np.mgrid[0:h, 0:w]
    out = {
        "zeros": np.zeros((h, w), np.uint8),
        "full": np.full((h, w), 255, np.uint8),
        "checker2": (((yy // 2 + xx // 2) & 1) * 255).astype(np.uint8),
        "checker1": (((yy + xx) & 1) * 255).astype(np.uint8),
        "noise": rng.integers(0, 256, size=(h, w)).astype(np.uint8),
        "lownoise": rng.integers(100, 112, size=(h, w)).astype(np.uint8),
    }
    lines = np.full((h, w), 128, np.uint8)
    lines[:, 0] = 255
    lines[:, w - 1] = 0
    lines[0, :] = 0
    lines[h - 1, :] = 255
    lines[20:h - 20:37, :] = 250
    lines[:, 20:w - 20:41] = 5
    out["lines"] = lines
    blobs = np.full((h, w), 30, np.uint8)
    for (cy, cx) in rng.integers(30, min(h, w) - 30, size=(40, 2)):
        blobs[cy - 2:cy + 3, cx - 2:cx + 3] = 220
    out["sparse_blobs"] = blobs
    return out


# ------------------------------------------------------------------------------------------------
# fr3_walking-shaped RGB-D stream: a textured wall seen by a translating / slightly rotating camera
# ------------------------------------------------------------------------------------------------
class WallStream:
    """Camera looking at the plane z_w = depth0 (world = first camera frame), moving on a Lissajous path
    with a small in-plane roll.  Pixel (u,v) of frame t back-projects to the plane exactly, so depth maps
    and ground-truth poses are exact (depth = z_c of the plane along each pixel ray)."""

    def __init__(self, seed: int = 1234, n: int = 827, depth0: float = 2.0, h: int = H, w: int = W,
                 tex_scale: float = 2.0):
        self.seed, self.n, self.depth0, self.h, self.w = seed, n, depth0, h, w
        self.th, self.tw = int(h * tex_scale) + 2 * 256, int(w * tex_scale) + 2 * 256
        self.tex = texture(seed, self.th, self.tw, nrect=1600, nline=800)
        self.ppm = FX / depth0   # texture pixels per metre at the wall for unit zoom

    def pose(self, t: int) -> np.ndarray:
        """Tcw (4x4 float32): world -> camera t."""
        a = 2 * np.pi * t / 240.0
        cx_w, cy_w, cz_w = 0.25 * np.sin(a), 0.12 * np.sin(2 * a), 0.15 * np.sin(0.5 * a)  # camera centre
        roll = np.deg2rad(2.0) * np.sin(3 * a)
        c, s = np.cos(roll), np.sin(roll)
        Rcw = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], np.float64)
        twc = np.array([cx_w, cy_w, cz_w])
        T = np.eye(4)
        T[:3, :3] = Rcw
        T[:3, 3] = -Rcw @ twc
        return T.astype(np.float32)

    def frame(self, t: int):
        """-> gray u8 (h,w), depth f32 metres (h,w), rgb u8 (h,w,3), Tcw f32 4x4."""
        T = self.pose(t).astype(np.float64)
        Rcw, tcw = T[:3, :3], T[:3, 3]
        Rwc = Rcw.T
        twc = -Rwc @ tcw
        v, u = np.mgrid[0:self.h, 0:self.w].astype(np.float64)
        rays = np.stack([(u - CX) / FX, (v - CY) / FY, np.ones_like(u)], -1) @ Rwc.T   # world-frame rays
        lam = (self.depth0 - twc[2]) / rays[..., 2]
        Xw = twc + rays * lam[..., None]
        depth = (lam * 1.0).astype(np.float32)   # z_c: ray has unit z in camera frame
        tx = Xw[..., 0] * self.ppm + self.tw / 2.0
        ty = Xw[..., 1] * self.ppm + self.th / 2.0
        x0 = np.clip(np.floor(tx).astype(int), 0, self.tw - 2)
        y0 = np.clip(np.floor(ty).astype(int), 0, self.th - 2)
        fx = np.clip(tx - x0, 0, 1)
        fy = np.clip(ty - y0, 0, 1)
        tex = self.tex
        val = (tex[y0, x0] * (1 - fx) + tex[y0, x0 + 1] * fx) * (1 - fy) + \
              (tex[y0 + 1, x0] * (1 - fx) + tex[y0 + 1, x0 + 1] * fx) * fy
        rng = np.random.Generator(np.random.PCG64(10 ** 6 + t))
        gray = np.clip(np.rint(val + rng.normal(0, 2.0, size=val.shape)), 0, 255).astype(np.uint8)
        rgb = np.stack([gray, np.clip(gray.astype(np.int32) + 10, 0, 255).astype(np.uint8),
                        (255 - gray)], -1).astype(np.uint8)
        # TUM-style quantisation: u16 = round(5000 z) then / 5000 in float (Tracking.cc:361-367)
        dq = np.rint(depth.astype(np.float64) * DEPTH_FACTOR).astype(np.uint16)
        depth = (dq.astype(np.float32) * np.float32(1.0 / DEPTH_FACTOR)).astype(np.float32)
        return gray, depth, rgb, self.pose(t)


# ------------------------------------------------------------------------------------------------
# Room box stream: textured planes of a 6 x 3 x 6 m room rendered by ray-plane intersection
# ------------------------------------------------------------------------------------------------
class RoomStream:
    """Camera inside a 6 m (x) x 3 m (y, pointing DOWN like the camera's y axis) x 6 m (z) room whose six faces carry
    textures.  The camera centre moves on a 0.5 m Lissajous around the room centre, 1.2 m above the floor, and pans
    (yaw) / nods (pitch) / rolls by at most ~1.3 deg per frame, so walls, floor and ceiling enter and leave the
    0.5 - 3.0 m gate of the occupancy path and the scene is not a single plane.  Every pixel's ray is intersected with
    the six planes exactly, so depth (z_c), the ground-truth pose and the GT floor mask (ground label of InsertScan's
    mode B, perfect/src/MapDrawer.cc:961-969) are exact."""

    HALF = np.array([3.0, 0.0, 3.0])
    Y_FLOOR, Y_CEIL = 1.2, -1.8

    def __init__(self, seed: int = 1234, n: int = 827, h: int = H, w: int = W, ppm: float = 230.0, period: float = 200.0):
        self.seed, self.n, self.h, self.w, self.ppm, self.period = seed, n, h, w, ppm, period
        ts = int(6.0 * ppm) + 8
        self.ts = ts
        # six faces: x = -3, x = +3, z = -3, z = +3 (walls, 6 x 3 m), floor, ceiling (6 x 6 m)
        self.tex = [texture(seed + 10 * k, int(3.0 * ppm) + 8 if k < 4 else ts, ts, nrect=1500 if k < 4 else 3000,
                            nline=700 if k < 4 else 1400) for k in range(6)]

    def pose(self, t: int) -> np.ndarray:
        a = 2 * np.pi * t / self.period
        twc = np.array([0.5 * np.sin(a), 0.08 * np.sin(2 * a), 0.5 * np.sin(1.5 * a + 0.4)])
        yaw = np.deg2rad(40.0) * np.sin(a) + 2 * np.pi * t / (8 * self.period)
        pitch = -(np.deg2rad(8.0) * np.sin(0.7 * a + 1.0) + np.deg2rad(12.0))   # y points down: negative = looking down at the floor
        roll = np.deg2rad(2.0) * np.sin(3 * a)
        cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        Rz = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
        Rwc = Ry @ Rx @ Rz
        T = np.eye(4)
        T[:3, :3] = Rwc.T
        T[:3, 3] = -Rwc.T @ twc
        return T.astype(np.float32)

    def frame(self, t: int, with_label: bool = False):
        """-> gray u8, depth f32 metres (TUM-quantised), rgb u8 (h,w,3), Tcw f32 4x4 [, floor label u8 (h,w)]."""
        T = self.pose(t).astype(np.float64)
        Rcw, tcw = T[:3, :3], T[:3, 3]
        Rwc = Rcw.T
        twc = -Rwc @ tcw
        v, u = np.mgrid[0:self.h, 0:self.w].astype(np.float64)
        rays = np.stack([(u - CX) / FX, (v - CY) / FY, np.ones_like(u)], -1) @ Rwc.T   # unit z_c per ray
        best = np.full((self.h, self.w), np.inf)
        face = np.zeros((self.h, self.w), np.int8)
        planes = [(0, -3.0), (0, 3.0), (2, -3.0), (2, 3.0), (1, self.Y_FLOOR), (1, self.Y_CEIL)]
        with np.errstate(divide="ignore", invalid="ignore"):
            for k, (ax, c) in enumerate(planes):
                lam = (c - twc[ax]) / rays[..., ax]
                ok = (lam > 1e-6) & (lam < best)
                best = np.where(ok, lam, best)
                face = np.where(ok, k, face).astype(np.int8)
        Xw = twc + rays * best[..., None]
        val = np.zeros((self.h, self.w))
        for k, (ax, c) in enumerate(planes):
            m = face == k
            if not m.any():
                continue
            P = Xw[m]
            if ax == 0:
                a_, b_ = P[:, 2] + 3.0, P[:, 1] - self.Y_CEIL
            elif ax == 2:
                a_, b_ = P[:, 0] + 3.0, P[:, 1] - self.Y_CEIL
            else:
                a_, b_ = P[:, 0] + 3.0, P[:, 2] + 3.0
            tex = self.tex[k]
            tx, ty = a_ * self.ppm + 2.0, b_ * self.ppm + 2.0
            x0 = np.clip(np.floor(tx).astype(int), 0, tex.shape[1] - 2)
            y0 = np.clip(np.floor(ty).astype(int), 0, tex.shape[0] - 2)
            fx_, fy_ = np.clip(tx - x0, 0, 1), np.clip(ty - y0, 0, 1)
            val[m] = (tex[y0, x0] * (1 - fx_) + tex[y0, x0 + 1] * fx_) * (1 - fy_) + \
                     (tex[y0 + 1, x0] * (1 - fx_) + tex[y0 + 1, x0 + 1] * fx_) * fy_
        rng = np.random.Generator(np.random.PCG64(10 ** 6 + t))
        gray = np.clip(np.rint(val + rng.normal(0, 2.0, size=val.shape)), 0, 255).astype(np.uint8)
        rgb = np.stack([gray, np.clip(gray.astype(np.int32) + 10, 0, 255).astype(np.uint8), (255 - gray)], -1).astype(np.uint8)
        dq = np.rint(np.minimum(best, 13.0) * DEPTH_FACTOR).astype(np.uint16)
        depth = (dq.astype(np.float32) * np.float32(1.0 / DEPTH_FACTOR)).astype(np.float32)
        if with_label:
            return gray, depth, rgb, self.pose(t), (face == 4).astype(np.uint8)
        return gray, depth, rgb, self.pose(t)


# ------------------------------------------------------------------------------------------------
# Dense flow fields for the dynamic-mask stage (perfect/src/Flow.cc): what calcOpticalFlowFarneback hands to pyrUp
# ------------------------------------------------------------------------------------------------
def flow_field(seed, rows, cols, blobs=6):
    """Smooth background flow (camera motion, |flow| ~ 2-5 px) + a few fast-moving blobs (|flow| up to ~15 px) + noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    f = np.zeros((rows, cols, 2), np.float32)
    f[..., 0] = 2.5 + 1.5 * np.sin(xx / 37.0) + rng.normal(0, 0.3, (rows, cols))
    f[..., 1] = -1.0 + 2.0 * np.cos(yy / 29.0) + rng.normal(0, 0.3, (rows, cols))
    for _ in range(blobs):
        cy, cx = rng.integers(0, rows), rng.integers(0, cols)
        ry, rx = rng.integers(4, max(rows // 4, 5)), rng.integers(4, max(cols // 4, 5))
        v = rng.normal(0, 9.0, 2).astype(np.float32)
        m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
        f[m] += v
    return f.astype(np.float32)
