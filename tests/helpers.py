"""Shared builders for the parity tests: frames -> matcher views, through the ORACLE (CPU) only."""
import numpy as np

from orb_slam2_ssd_semantic_b200 import synth
from orb_slam2_ssd_semantic_b200._abi import FrameView, LastView


def frame_views(oracle, K_cur, D_cur, depth_cur, T_cur, K_last, D_last, depth_last, T_last, sf, obs=1,
                rows=480, cols=640):
    """What Tracking hands to SearchByProjection(Cur, Last): Cur = keypoints + uRight; Last = every keypoint with
    depth unprojected into a 'MapPoint' carrying its own descriptor."""
    ur_c, _, _, _ = oracle.stereo_unproject(K_cur, depth_cur, T_cur, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF)
    cur = FrameView(K_cur["x"], K_cur["y"], K_cur["octave"], K_cur["angle"], ur_c, D_cur, T_cur, synth.FX, synth.FY,
                    synth.CX, synth.CY, synth.BF, 0.0, float(cols), 0.0, float(rows), sf)
    _, _, xw, valid = oracle.stereo_unproject(K_last, depth_last, T_last, synth.FX, synth.FY, synth.CX, synth.CY,
                                              synth.BF)
    last = LastView(xw, valid, K_last["octave"], K_last["angle"], D_last, T_last,
                    mp_obs=np.full(len(K_last), obs, np.int32))
    return cur, last
