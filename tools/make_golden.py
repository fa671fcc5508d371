#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz.  Runs ONLY in the build container: needs cv2 (the real OpenCV primitives) and
/root/reference (map.bin, octomap.ot).  The fixtures are small and committed; tests never read /root/reference.

  extract_*.npz   image + keypoints + descriptors from oracle/orb_cv2.py (cv2 4.13 primitives, Python restatement)
  match_mapbin_*.npz  two real keyframes of the reference's map.bin (real ORB keypoints/descriptors/map points,
                      intrinsics of perfect/Examples/RGB-D/my_rgbd_ty_api_adj.yaml) + the match vector computed by
                      the pure-Python restatement oracle/match_py.py
  octomap_logodds.npz  distinct node log-odds values of the reference's octomap.ot
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.orb_cv2 import ORBextractorCV2  # noqa: E402
from oracle import match_py  # noqa: E402
from orb_slam2_ssd_semantic_b200 import synth  # noqa: E402
from orb_slam2_ssd_semantic_b200._abi import FrameView, LastView  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
os.makedirs(G, exist_ok=True)


def extract_fixtures():
    import cv2
    cases = {"synth0": (synth.synth_frame(1234, 0), (1000, 1.2, 8, 20, 7)),
             "lines": (synth.adversarial_frames()["lines"], (1000, 1.2, 8, 20, 7)),
             "checker1_qvga": (synth.adversarial_frames(240, 320)["checker1"], (500, 1.2, 8, 20, 7)),
             "synth_small_2000": (synth.synth_frame(7, 2, h=300, w=400), (2000, 1.2, 6, 20, 7))}
    for name, (img, prm) in cases.items():
        E = ORBextractorCV2(*prm)
        K, D = E(img)
        np.savez_compressed(os.path.join(G, "extract_%s.npz" % name), image=img, params=np.array(prm, np.float64),
                            kps=K, desc=D, candidates=np.array(E.candidates_per_level, np.int32),
                            cv2_version=np.array(cv2.__version__))
        print(name, len(K))


def load_mapbin():
    b = open("/root/reference/map.bin", "rb").read()
    off = 0
    nmp, = struct.unpack_from("<Q", b, off); off += 8
    mp = np.frombuffer(b, dtype=np.dtype([("id", "<u8"), ("p", "<f4", 3)]), count=nmp, offset=off); off += nmp * 20
    nkf, = struct.unpack_from("<Q", b, off); off += 8
    kpdt = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("resp", "<f4"), ("oct", "<i4"),
                     ("desc", "u1", 32), ("mp", "<u8")])
    kfs = []
    for _ in range(nkf):
        kid, ts = struct.unpack_from("<Qd", b, off); off += 16
        t = np.frombuffer(b, "<f4", 3, off); off += 12
        q = np.frombuffer(b, "<f4", 4, off); off += 16
        N, = struct.unpack_from("<i", b, off); off += 4
        k = np.frombuffer(b, kpdt, N, off); off += N * kpdt.itemsize
        kfs.append((kid, ts, t, q, k))
    return mp, kfs


def q2T(q, t):
    x, y, z, w = [float(v) for v in q]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T.astype(np.float32)


def match_fixtures():
    mp, kfs = load_mapbin()
    by_id = {kf[0]: kf for kf in kfs}
    fx, fy, cx, cy, bf = 558.957, 559.094, 306.279, 268.493, 27.95   # my_rgbd_ty_api_adj.yaml
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    sf = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2), np.float32)])).astype(np.float32)
    pairs = [(7, 8), (20, 21), (100, 101)]
    for a, b_ in pairs:
        if a not in by_id or b_ not in by_id:
            continue
        _, _, ta, qa, ka = by_id[a]
        _, _, tb, qb, kb = by_id[b_]
        Ta, Tb = q2T(qa, ta), q2T(qb, tb)
        has = ka["mp"] != 2 ** 64 - 1
        xw = np.zeros((len(ka), 3), np.float32)
        xw[has] = mp["p"][ka["mp"][has].astype(np.int64)]
        # uRight: unknown (no depth in map.bin) -> -1 for all (mono-like keypoints)
        cur = FrameView(kb["x"], kb["y"], kb["oct"], kb["angle"], np.full(len(kb), -1, np.float32), kb["desc"], Tb, fx, fy,
                        cx, cy, bf, 0.0, 640.0, 0.0, 480.0, sf)
        last = LastView(xw, has.astype(np.uint8), ka["oct"], ka["angle"], ka["desc"], Ta, mp_obs=np.ones(len(ka), np.int32))
        for th in (15.0,):
            n, m = match_py.search_by_projection_last(cur, last, th, False, True)
            print("mapbin pair", a, b_, "matches", n)
            np.savez_compressed(os.path.join(G, "match_mapbin_%d_%d.npz" % (a, b_)), cur_x=cur.x, cur_y=cur.y,
                                cur_oct=cur.octave, cur_angle=cur.angle, cur_uright=cur.uright, cur_desc=cur.desc,
                                cur_Tcw=cur.Tcw, last_xw=last.xw, last_valid=last.valid, last_oct=last.octave,
                                last_angle=last.angle, last_desc=last.mp_desc, last_Tcw=last.Tcw,
                                cam=np.array([fx, fy, cx, cy, bf], np.float32), sf=sf, th=np.float32(th),
                                nmatches=np.int32(n), cur2last=m)


def octomap_fixture():
    b = open("/root/reference/octomap.ot", "rb").read()
    i = b.index(b"data\n") + 5
    header = b[:i].decode("ascii", "replace")
    size = int([l for l in header.splitlines() if l.startswith("size")][0].split()[1])
    res = float([l for l in header.splitlines() if l.startswith("res")][0].split()[1])
    body = np.frombuffer(b[i:i + size * 8], dtype=np.dtype([("v", "<f4"), ("rgb", "u1", 3), ("child", "u1")]))
    vals = np.unique(body["v"])
    leaf_vals = np.unique(body["v"][body["child"] == 0])
    print("octomap nodes", size, "res", res, "distinct values", len(vals), "leaf distinct", len(leaf_vals))
    np.savez_compressed(os.path.join(G, "octomap_logodds.npz"), values=vals, leaf_values=leaf_vals, size=np.int64(size),
                        res=np.float64(res))
    # the whole node array (pre-order) for the .ot exporter round trip (orb_slam2_ssd_semantic_b200/octree_io.py)
    np.savez_compressed(os.path.join(G, "octomap_nodes.npz"), v=body["v"], rgb=body["rgb"], child=body["child"],
                        header=np.array(header))


if __name__ == "__main__":
    extract_fixtures()
    match_fixtures()
    octomap_fixture()
