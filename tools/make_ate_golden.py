"""Builds tests/golden/ate_f3_walking.npz from the reference's trajectory fixtures (tool/groundtruth.txt = TUM
fr3_walking_xyz ground truth, tool/src.txt = the reference's own estimated trajectory) and the figures README.md:156-163
prints for them.  Run in the container where /root/reference exists."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_ssd_semantic_b200 import ate  # noqa: E402

REF = "/root/reference/tool"
gt = ate.read_trajectory(os.path.join(REF, "groundtruth.txt"))
est = ate.read_trajectory(os.path.join(REF, "src.txt"))
gs, es = sorted(gt), sorted(est)
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ate_f3_walking.npz"),
                    gt_stamp=np.array(gs), gt_xyz=np.array([gt[s][:3] for s in gs]), est_stamp=np.array(es),
                    est_xyz=np.array([est[s][:3] for s in es]),
                    readme=np.array([826, 0.702233, 0.582247, 0.520550, 0.392580, 0.077351, 1.476973]))   # README.md:157-163
r = ate.evaluate(gt, est)
print({k: v for k, v in r.items() if k not in ("rot", "trans")})
