"""configs[4] pieces that need no GPU: the ATE evaluator pinned by the reference's own trajectory fixtures, and the reduced
tracking loop run on the CPU oracle."""
import os

import numpy as np

from orb_slam2_ssd_semantic_b200 import ate, synth
from orb_slam2_ssd_semantic_b200.tracking import ReducedTracker

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_ate_evaluator_reproduces_reference_readme():
    """tool/src.txt vs tool/groundtruth.txt of the reference (committed as tests/golden/ate_f3_walking.npz by
    tools/make_ate_golden.py): the evaluator must print what README.md:156-163 prints, to the 6 digits shown there."""
    z = np.load(os.path.join(G, "ate_f3_walking.npz"))
    gt = {float(s): list(p) for s, p in zip(z["gt_stamp"], z["gt_xyz"])}
    est = {float(s): list(p) for s, p in zip(z["est_stamp"], z["est_xyz"])}
    r = ate.evaluate(gt, est)
    pairs, rmse, mean, median, std, mn, mx = z["readme"]
    assert r["compared_pose_pairs"] == int(pairs)
    for got, want in ((r["rmse"], rmse), (r["mean"], mean), (r["median"], median), (r["std"], std), (r["min"], mn), (r["max"], mx)):
        assert abs(got - want) < 5e-7, (got, want)


def test_ate_of_a_rigidly_moved_trajectory_is_zero():
    rng = np.random.default_rng(0)
    P = rng.normal(0, 1, (200, 3))
    a = 0.7
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    Q = P @ R.T + np.array([3.0, -2.0, 0.5])
    st = np.arange(200) * 0.033
    r = ate.evaluate({s: p for s, p in zip(st, Q)}, {s + 0.004: p for s, p in zip(st, P)})
    assert r["compared_pose_pairs"] == 200 and r["rmse"] < 1e-12


def run_loop(backend_extract, backend_unproject, backend_match, nframes, seed=21, step=2):
    rs = synth.RoomStream(seed=seed, n=nframes * step + 1)
    sf = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2), np.float32)])).astype(np.float32)
    trk = ReducedTracker(backend_extract, backend_unproject, backend_match, sf)
    gt = {}
    for i in range(nframes):
        gray, depth, rgb, T = rs.frame(i * step)
        trk.track(gray, depth, T_init=T if i == 0 else None)
        gt[i / 30.0] = (-T[:3, :3].astype(np.float64).T @ T[:3, 3].astype(np.float64)).tolist()
    return trk, gt


def test_reduced_tracking_loop_on_the_oracle(oracle):
    R = oracle.RefExtractor(1000, 1.2, 8, 20, 7)
    unproj = lambda K, d, T: oracle.stereo_unproject(K, d, T, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF)
    match = lambda cur, last, th: oracle.search_by_projection_last(cur, last, th, False, 0.9, True)
    trk, gt = run_loop(R, unproj, match, 8)
    assert min(trk.nmatches[1:]) > 100
    r = ate.evaluate(gt, trk.trajectory([i / 30.0 for i in range(8)]))
    assert r["compared_pose_pairs"] == 8 and r["rmse"] < 0.02
