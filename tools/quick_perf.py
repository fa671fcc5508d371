"""Quick per-stage timing of the stream pipeline on the GPU (development aid, not the bench)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from orb_slam2_ssd_semantic_b200 import StreamTracker, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nfeat = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ws = synth.WallStream(seed=1234, n=F)
t0 = time.time()
fr = [ws.frame(t) for t in range(F)]
print("gen %.1fs" % (time.time() - t0))
gray = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
depth = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
T = torch.from_numpy(np.stack([f[3] for f in fr])).cuda()
st = StreamTracker(nfeat, 1.2, 8, 20, 7, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, max_frames=F)
for _ in range(3):
    st.track_batch_device(gray.data_ptr(), depth.data_ptr(), T.data_ptr(), F, 480, 640)
st.sync()
st.profile_enable(True)
K = 5
t0 = time.time()
for _ in range(K):
    st.track_batch_device(gray.data_ptr(), depth.data_ptr(), T.data_ptr(), F, 480, 640)
st.sync()
dt = time.time() - t0
ms, frames, runs = st.profile_read()
print("wall: %.3f ms/batch, %.1f frames/s" % (dt / K * 1e3, F * K / dt))
tot = sum(ms.values())
for k, v in ms.items():
    print("  %-12s %8.3f ms/batch  %5.1f%%  %7.2f us/frame" % (k, v / runs, 100 * v / tot, v / frames * 1e3))
print("  total        %8.3f ms/batch -> %.0f frames/s" % (tot / runs, frames / tot * 1e3))
