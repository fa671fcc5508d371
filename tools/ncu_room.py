"""Workload for ncu captures: a batch of room-stream frames through the tracker + the labelled keyframe inserts, three
times (development aid; the numbers a profiler run prints are never bench values).  The last iteration sits inside
cudaProfilerStart/Stop, so `ncu --profile-from-start off` captures exactly one warm pass.
usage: ncu_room.py [frames] [nfeatures] [iterations]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from orb_slam2_ssd_semantic_b200 import PointCloudMapping, StreamTracker, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 96
nfeat = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rs = synth.RoomStream(seed=1234, n=F)
fr = [rs.frame(t, with_label=True) for t in range(F)]
gray = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
depth = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
rgb = torch.from_numpy(np.stack([f[2] for f in fr])).cuda()
lab = torch.from_numpy(np.stack([f[4] for f in fr])).cuda()
Th = np.stack([f[3] for f in fr]).astype(np.float32)
T = torch.from_numpy(Th).cuda()
st = StreamTracker(nfeat, 1.2, 8, 20, 7, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, max_frames=F)
pcm = PointCloudMapping(0.05)
kfs = list(range(0, F, 12))
st.profile_enable(True)
for it in range(ITERS):
    if it == ITERS - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    st.track_batch_device(gray.data_ptr(), depth.data_ptr(), T.data_ptr(), F, 480, 640)
    st.sync()
    pcm.insert_keyframes_device(depth.data_ptr(), rgb.data_ptr(), 480, 640, kfs, Th[kfs], synth.FX, synth.FY, synth.CX, synth.CY,
                                d_label=lab.data_ptr())
    pcm.sync()
torch.cuda.profiler.stop()
ms, frames, runs = st.profile_read()
print({k: round(v / runs, 3) for k, v in ms.items()}, "ms per %d-frame batch" % F)
