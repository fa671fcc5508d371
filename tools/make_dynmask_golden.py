#!/usr/bin/env python3
"""Writes tests/golden/dynmask_*.npz from cv2 (4.13 in this image): the third-party code perfect/src/Flow.cc calls after
its optical flow -- pyrUp, erode, erode, dilate with getStructuringElement(MORPH_ELLIPSE, 21x21) -- run on seeded flow
fields, so that the oracle (oracle/dynmask_py.py) and the GPU path are checked against OpenCV's own outputs on boxes
without cv2.  usage: python tools/make_dynmask_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


from orb_slam2_ssd_semantic_b200.synth import flow_field  # noqa: E402


def cv_mask(flow, thr, shape=None):
    flow2 = cv2.pyrUp(flow, dstsize=(flow.shape[1] * 2, flow.shape[0] * 2))
    thr = max(np.float32(thr), np.float32(40.0))
    t2 = flow2[..., 0] * flow2[..., 0] + flow2[..., 1] * flow2[..., 1]
    mask = np.ones(flow2.shape[:2] if shape is None else shape, np.uint8)        # Flow.cc:25
    mask[:flow2.shape[0], :flow2.shape[1]] = (t2 < thr).astype(np.uint8)         # :31-41
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (21, 21), (10, 10))
    m = mask.copy()
    m = cv2.erode(m, k)
    m = cv2.erode(m, k)
    m = cv2.dilate(m, k)
    return flow2, mask, m, k


def main():
    out = os.path.join(ROOT, "tests", "golden")
    cases = [("a", 11, 120, 160, 40.0), ("b", 12, 60, 80, 64.0), ("c", 13, 33, 47, 10.0), ("d", 14, 2, 2, 40.0)]
    cases.append(("e", 15, 16, 23, 40.0))          # odd-sized gray image: 33 x 47
    for name, seed, rows, cols, thr in cases:
        flow = flow_field(seed, rows, cols)
        shape = (33, 47) if name == "e" else (2 * rows, 2 * cols)
        flow2, m0, m, k = cv_mask(flow, thr, shape)
        np.savez_compressed(os.path.join(out, "dynmask_%s.npz" % name), flow=flow, thr=np.float32(thr),
                            flow2_sample=flow2[::7, ::5].copy(), flow2_border=np.concatenate([flow2[0].ravel(), flow2[-1].ravel(),
                                                                                              flow2[:, 0].ravel(), flow2[:, -1].ravel()]),
                            mask0=np.packbits(m0), mask=np.packbits(m), shape=np.array(shape), element=k, cv2_version=cv2.__version__)
        print(name, rows, cols, thr, "static fraction %.3f -> %.3f" % (m0.mean(), m.mean()))


if __name__ == "__main__":
    main()
