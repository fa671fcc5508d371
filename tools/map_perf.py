"""Standalone timing of the keyframe -> occupancy path (development aid): 22 keyframes per step like bench.py."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from orb_slam2_ssd_semantic_b200 import PointCloudMapping, synth

F, KF = 256, 12
ws = synth.WallStream(seed=1234, n=F)
kfs = list(range(0, F, KF))
fr = [ws.frame(t) for t in kfs]
depth = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
rgb = torch.from_numpy(np.stack([f[2] for f in fr])).cuda()
T = np.stack([f[3] for f in fr]).astype(np.float32)
pcm = PointCloudMapping(0.05)
idx = list(range(len(kfs)))
for _ in range(3):
    pcm.insert_keyframes_device(depth.data_ptr(), rgb.data_ptr(), 480, 640, idx, T, synth.FX, synth.FY, synth.CX, synth.CY)
pcm.sync()
ext = torch.cuda.ExternalStream(pcm.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record(ext)
for _ in range(K):
    pcm.insert_keyframes_device(depth.data_ptr(), rgb.data_ptr(), 480, 640, idx, T, synth.FX, synth.FY, synth.CX, synth.CY)
e1.record(ext)
pcm.sync()
print("mapping alone: %.3f ms per %d keyframes (%.1f us/keyframe), leaves %d" %
      (e0.elapsed_time(e1) / K, len(kfs), e0.elapsed_time(e1) / K / len(kfs) * 1e3, pcm.num_leaves()))
