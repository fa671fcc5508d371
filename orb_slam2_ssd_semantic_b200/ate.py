"""Absolute trajectory error of an estimated camera trajectory against ground truth, the way the reference evaluates its
runs (tool/evaluate_ate.py + tool/associate.py, the TUM RGB-D benchmark tools; README.md:140-163 quotes their output):
time-stamp association (greedy, best time difference first, within max_difference), rigid alignment by Horn's closed form
(SVD of the cross-covariance, reflection guarded), per-pair translational error statistics.

Host-side evaluation code of BASELINE.json configs[4] ("ATE vs reference CPU on identical inputs"): pinned by the
reference's own fixtures (tests/golden/ate_f3_walking.npz made from tool/src.txt / tool/groundtruth.txt) to the six
figures the README prints for them."""
from __future__ import annotations

import numpy as np


def read_trajectory(path: str) -> dict:
    """'stamp tx ty tz [qx qy qz qw]' lines (comments '#', separators space / comma / tab) -> {stamp: [floats]}."""
    out = {}
    with open(path) as f:
        for line in f.read().replace(",", " ").replace("\t", " ").split("\n"):
            if not line or line[0] == "#":
                continue
            v = [t for t in line.split(" ") if t.strip() != ""]
            if len(v) > 1:
                out[float(v[0])] = [float(t) for t in v[1:]]
    return out


def associate(first_stamps, second_stamps, offset: float = 0.0, max_difference: float = 0.02):
    """Greedy one-to-one association of two stamp lists: candidate pairs closer than max_difference, taken in order of
    (difference, first stamp, second stamp); returns the matches sorted by first stamp."""
    a = np.asarray(sorted(first_stamps), np.float64)
    b = np.asarray(sorted(second_stamps), np.float64)
    cand = []
    j0 = 0
    for x in a:
        while j0 < len(b) and b[j0] + offset <= x - max_difference:
            j0 += 1
        j = j0
        while j < len(b) and b[j] + offset < x + max_difference:
            d = abs(x - (b[j] + offset))
            if d < max_difference:
                cand.append((d, float(x), float(b[j])))
            j += 1
    cand.sort()
    used_a, used_b, matches = set(), set(), []
    for _, x, y in cand:
        if x not in used_a and y not in used_b:
            used_a.add(x)
            used_b.add(y)
            matches.append((x, y))
    matches.sort()
    return matches


def align(model: np.ndarray, data: np.ndarray):
    """Horn alignment of two 3 x n point sets: rot, trans with rot @ model + trans ~ data, and the per-point error."""
    model = np.asarray(model, np.float64)
    data = np.asarray(data, np.float64)
    mm, dm = model.mean(1, keepdims=True), data.mean(1, keepdims=True)
    mz, dz = model - mm, data - dm
    W = np.zeros((3, 3))
    for c in range(model.shape[1]):
        W += np.outer(mz[:, c], dz[:, c])
    U, d, Vh = np.linalg.svd(W.T)
    S = np.identity(3)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0:
        S[2, 2] = -1
    rot = U @ S @ Vh
    trans = dm - rot @ mm
    err = rot @ model + trans - data
    return rot, trans, np.sqrt((err * err).sum(0))


def evaluate(gt: dict, est: dict, offset: float = 0.0, scale: float = 1.0, max_difference: float = 0.02) -> dict:
    """gt / est: {stamp: [tx, ty, tz, ...]} -> the figures evaluate_ate.py --verbose prints."""
    matches = associate(list(gt.keys()), list(est.keys()), offset, max_difference)
    if len(matches) < 2:
        raise ValueError("Couldn't find matching timestamp pairs between groundtruth and estimated trajectory")
    first = np.array([[float(v) for v in gt[a][0:3]] for a, _ in matches]).T
    second = np.array([[float(v) * float(scale) for v in est[b][0:3]] for _, b in matches]).T
    rot, trans, err = align(second, first)
    return {"compared_pose_pairs": len(err), "rmse": float(np.sqrt(np.dot(err, err) / len(err))), "mean": float(np.mean(err)),
            "median": float(np.median(err)), "std": float(np.std(err)), "min": float(np.min(err)), "max": float(np.max(err)),
            "rot": rot, "trans": trans}
