"""CPU restatement (numpy) of the dynamic-mask stages -- TEST INFRASTRUCTURE ONLY (tests/, smoke): never imported by the
product.  Pinned against the third-party code the reference calls: cv2 4.13's pyrUp / erode / dilate /
getStructuringElement (tools/make_dynmask_golden.py wrote tests/golden/dynmask_*.npz from cv2's outputs;
tests/test_dynmask_cpu.py checks this file against them and, where cv2 is importable, against cv2 live); filter_keypoints is
pinned to the reference's own masked Frame constructor (perfect/src/Frame.cc compiled unmodified: oracle/_ref/librefperfect.so).

  perfect/src/Flow.cc:30      pyrUp(flow, flow2, Size(2 cols, 2 rows))        -> pyr_up
  perfect/src/Flow.cc:24,31-41 threshold loop                                  -> flow_mask
  perfect/src/Flow.cc:42-47   getStructuringElement + erode, erode, dilate    -> ellipse, morph, mask_from_flow
  perfect/src/Frame.cc:356-377 masked Frame constructor's keypoint loop        -> filter_keypoints
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def pyr_up(src: np.ndarray) -> np.ndarray:
    """cv::pyrUp of a float image (rows x cols x cn) to exactly twice the size: OpenCV's pyrUp_<FltCast<float,6>> float
    sequence -- row pass x*6 + left + right (even) / (x + right)*4 (odd), reflect-101 on the left (x*6 + next*2),
    replicate on the right (prev + x*7, x*8); column pass the same with rows borderInterpolate(2y, 2h, REFLECT_101)/2,
    times 1/64."""
    src = np.asarray(src, np.float32)
    h, w = src.shape[:2]
    assert h >= 2 and w >= 2
    rows = np.zeros((h, 2 * w) + src.shape[2:], np.float32)
    x0 = src
    rows[:, 0] = x0[:, 0] * F32(6) + x0[:, 1] * F32(2)
    rows[:, 1] = (x0[:, 0] + x0[:, 1]) * F32(4)
    if w > 2:
        mid = x0[:, 1:w - 1]
        rows[:, 2:2 * w - 2:2] = (mid * F32(6) + x0[:, 0:w - 2]) + x0[:, 2:w]
        rows[:, 3:2 * w - 2:2] = (mid + x0[:, 2:w]) * F32(4)
    rows[:, 2 * w - 2] = x0[:, w - 2] + x0[:, w - 1] * F32(7)
    rows[:, 2 * w - 1] = x0[:, w - 1] * F32(8)
    out = np.zeros((2 * h, 2 * w) + src.shape[2:], np.float32)
    s = F32(1.0 / 64.0)
    for y in range(h):
        yp = y - 1 if y > 0 else 1
        yn = y + 1 if y + 1 < h else h - 1
        r0, r1, r2 = rows[yp], rows[y], rows[yn]
        out[2 * y] = ((r1 * F32(6) + r0) + r2) * s
        out[2 * y + 1] = ((r1 + r2) * F32(4)) * s
    return out


def flow_mask(flow2: np.ndarray, binary_threshold: float) -> np.ndarray:
    """mask = 1, 0 where x*x + y*y >= max(threshold, 40) in float arithmetic (NaN compares false: 0)."""
    thr = F32(binary_threshold)
    if thr < F32(40.0):
        thr = F32(40.0)
    t2 = flow2[..., 0] * flow2[..., 0] + flow2[..., 1] * flow2[..., 1]
    with np.errstate(invalid="ignore"):
        return (t2 < thr).astype(np.uint8)


def ellipse(ksize: int = 21) -> np.ndarray:
    """cv::getStructuringElement(MORPH_ELLIPSE, Size(ksize, ksize)) with the default (centre) anchor."""
    r = c = ksize // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    el = np.zeros((ksize, ksize), np.uint8)
    for i in range(ksize):
        dy = i - r
        dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
        el[i, max(c - dx, 0):min(c + dx + 1, ksize)] = 1
    return el


def morph(m: np.ndarray, el: np.ndarray, erode: bool) -> np.ndarray:
    """cv::erode / cv::dilate, anchor at the centre, default border value (outside pixels never win)."""
    H, W = m.shape
    r = el.shape[0] // 2
    fill = 255 if erode else 0
    pad = np.full((H + 2 * r, W + 2 * r), fill, np.uint8)
    pad[r:r + H, r:r + W] = m
    out = np.full((H, W), fill, np.uint8)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            if el[dy + r, dx + r]:
                s = pad[r + dy:r + dy + H, r + dx:r + dx + W]
                out = np.minimum(out, s) if erode else np.maximum(out, s)
    return out


def mask_from_flow(flow: np.ndarray, binary_threshold: float = 40.0, shape=None) -> np.ndarray:
    """Flow::ComputeMask after calcOpticalFlowFarneback: half-resolution flow -> static / dynamic mask of the gray image's
    `shape` (default 2 rows x 2 cols); pixels the up-sampled flow does not cover keep the initial 1 (:25)."""
    el = ellipse(21)
    f2 = pyr_up(flow)
    shape = f2.shape[:2] if shape is None else tuple(shape)
    m = np.ones(shape, np.uint8)
    m[:f2.shape[0], :f2.shape[1]] = flow_mask(f2, binary_threshold)
    return morph(morph(morph(m, el, True), el, True), el, False)


def filter_keypoints(mask: np.ndarray, kps: np.ndarray, desc: np.ndarray):
    """perfect/src/Frame.cc:356-377: cv::sum(mask) > rows*cols*0.65 -> keep keypoints with mask.at<uchar>(pt.y, pt.x) == 1."""
    s = float(mask.astype(np.float64).sum())
    if not (s > mask.shape[0] * mask.shape[1] * 0.65):
        return kps.copy(), desc.copy()
    yy = kps["y"].astype(np.int32)      # float -> int conversion truncates (values are non-negative)
    xx = kps["x"].astype(np.int32)
    keep = mask[yy, xx] == 1
    return kps[keep].copy(), desc[keep].copy()
