#!/usr/bin/env python3
"""Launch-shape / formulation A-B of the extractor on one room-stream batch in ONE process (development aid; the timings
are CUDA-event stage times of StreamTracker, not bench values): for every configuration of b200orb_set_tuning() it
(1) runs the batch through the host-buffer call and compares every output byte with the baseline configuration,
(2) times the device-resident call per stage.  Writes gpurun_out/tune.json incrementally.
usage: tune_extractor.py [frames] [nfeatures]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
from orb_slam2_ssd_semantic_b200 import StreamTracker, _lib, synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 96
nfeat = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
OUT = os.path.join("gpurun_out", "tune.json")
os.makedirs("gpurun_out", exist_ok=True)
L = _lib.lib()
rs = synth.RoomStream(seed=1234, n=F)
fr = [rs.frame(t) for t in range(F)]
gray, depth, Th = np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr]), np.stack([f[3] for f in fr]).astype(np.float32)
d_gray, d_depth, d_T = torch.from_numpy(gray).cuda(), torch.from_numpy(depth).cuda(), torch.from_numpy(Th).cuda()
st = StreamTracker(nfeat, 1.2, 8, 20, 7, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, max_frames=F)
PAR_F = min(F, 24)   # frames of the byte-for-byte comparison


def outputs():
    kps, desc, nkp, c2l, nm = st.track_batch(gray[:PAR_F], depth[:PAR_F], Th[:PAR_F])
    return [a.copy() for a in (kps.view(np.uint8), desc, nkp, c2l, nm)]


def timed(reps=4):
    for _ in range(2):
        st.track_batch_device(d_gray.data_ptr(), d_depth.data_ptr(), d_T.data_ptr(), F, 480, 640)
    st.sync()
    st.profile_enable(True)
    st.profile_read()
    for _ in range(reps):
        st.track_batch_device(d_gray.data_ptr(), d_depth.data_ptr(), d_T.data_ptr(), F, 480, 640)
    st.sync()
    ms, frames, runs = st.profile_read()
    st.profile_enable(False)
    return {k: v / runs for k, v in ms.items()}


def same(a, b):
    n = a[2]
    if not (a[2] == b[2]).all() or not (a[4] == b[4]).all():
        return False
    cap = a[1].shape[1]
    for f in range(len(n)):
        if a[0].reshape(len(n), cap, 28)[f, :n[f]].tobytes() != b[0].reshape(len(n), cap, 28)[f, :n[f]].tobytes():
            return False
        if not (a[1][f, :n[f]] == b[1][f, :n[f]]).all() or not (a[3][f, :n[f]] == b[3][f, :n[f]]).all():
            return False
    return True


res = {"frames": F, "nfeatures": nfeat, "configs": []}
base = None


def run(mask, wpc, minb):
    global base
    _lib.check(L.b200orb_set_tuning(mask, wpc, minb))
    o = outputs()
    if base is None:
        base = o
    ms = timed()
    e = {"exp_mask": mask, "fast_wpc": wpc, "qt_minb": minb, "same_as_base": bool(same(base, o)),
         "ms": {k: round(v, 4) for k, v in ms.items()}, "total_ms": round(sum(ms.values()), 4)}
    res["configs"].append(e)
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)
    print(e, flush=True)
    return e


b = run(0, 8, 2)
w = {8: b}
for wpc in (4, 2, 1):
    w[wpc] = run(0, wpc, 2)
q = {2: b}
for minb in (3, 4):
    q[minb] = run(0, 8, minb)
best_wpc = min((k for k in w if w[k]["same_as_base"]), key=lambda k: w[k]["ms"]["fast"])
best_minb = min((k for k in q if q[k]["same_as_base"]), key=lambda k: q[k]["ms"]["quadtree"])
res["best"] = {"fast_wpc": best_wpc, "qt_minb": best_minb, "exp_mask": 0}
ok_mask = 0
for bit in (1, 2):                      # the formulations behind the experimental switch, one at a time
    e = run(bit, 8, 2)
    stage = "orient_desc" if bit == 1 else "fast"
    if e["same_as_base"] and e["ms"][stage] < b["ms"][stage]:
        ok_mask |= bit
res["best"]["exp_mask"] = ok_mask
res["best"]["exp_evaluated"] = True
e = run(ok_mask, best_wpc, best_minb)
res["best"]["total_ms"] = e["total_ms"]
res["best"]["same_as_base"] = e["same_as_base"]
res["best"]["speedup_vs_base"] = b["total_ms"] / e["total_ms"]
with open(OUT, "w") as f:
    json.dump(res, f, indent=1)
print("BEST", res["best"])
