"""configs[4] on the GPU: the reduced tracking loop (extract -> stereo-from-depth -> SearchByProjection with the motion
model -> pose update) driven by the B200 library gives the SAME trajectory, pose for pose, as the same loop driven by the
CPU oracle -- hence the same ATE on identical inputs."""
import numpy as np
import pytest

from orb_slam2_ssd_semantic_b200 import ate, synth
from tests.test_tracking_cpu import run_loop

pytestmark = pytest.mark.gpu


def test_reduced_tracking_loop_gpu_equals_cpu(oracle):
    from orb_slam2_ssd_semantic_b200 import ORBextractor, ORBmatcher, StreamTracker
    n = 40
    ex, m = ORBextractor(1000, 1.2, 8, 20, 7), ORBmatcher(0.9, True)
    st = StreamTracker(1000, 1.2, 8, 20, 7, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, 15.0, 0.9, True, 2)
    img_of, keep = {}, []

    def extract2(gray):
        K, D = ex(gray)
        img_of[id(K)] = gray
        keep.append(K)          # keeps id(K) unique for the whole run
        return K, D

    def unproject(K, depth, T):
        # the frame glue kernel lives behind the stream pipeline: one-frame batch (it re-extracts: same keypoints)
        kps, desc, nkp, c2l, nm = st.track_batch(img_of[id(K)][None], depth[None], np.asarray(T, np.float32)[None])
        assert nkp[0] == len(K) and kps[0, :len(K)].tobytes() == K.tobytes()
        return st.frame_glue(0, len(K))

    match = lambda cur, last, th: m.SearchByProjection(cur, last, th)
    gpu, gt = run_loop(extract2, unproject, match, n)
    R = oracle.RefExtractor(1000, 1.2, 8, 20, 7)
    cpu, _ = run_loop(R, lambda K, d, T: oracle.stereo_unproject(K, d, T, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF),
                      lambda cur, last, th: oracle.search_by_projection_last(cur, last, th, False, 0.9, True), n)
    assert gpu.nmatches == cpu.nmatches and min(gpu.nmatches[1:]) > 100
    for a, b in zip(gpu.poses, cpu.poses):
        assert a.tobytes() == b.tobytes()
    st_ = [i / 30.0 for i in range(n)]
    ra, rb = ate.evaluate(gt, gpu.trajectory(st_)), ate.evaluate(gt, cpu.trajectory(st_))
    assert ra["rmse"] == rb["rmse"] and ra["rmse"] < 0.03
