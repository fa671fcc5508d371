"""Phase cycle counters of k_match_last_fused (development aid).  Needs the instrumented build:
   make -C orb_slam2_ssd_semantic_b200/csrc timing && B200ORB_LIB=$PWD/orb_slam2_ssd_semantic_b200/libb200orb_timing.so python tools/match_timing.py"""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from orb_slam2_ssd_semantic_b200 import StreamTracker, synth, _lib

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ws = synth.WallStream(seed=1234, n=F)
fr = [ws.frame(t) for t in range(F)]
gray = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
depth = torch.from_numpy(np.stack([f[1] for f in fr])).cuda()
T = torch.from_numpy(np.stack([f[3] for f in fr])).cuda()
st = StreamTracker(1000, 1.2, 8, 20, 7, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, max_frames=F)
L = _lib.lib()
out = (C.c_ulonglong * 16)()
for it in range(3):
    st.track_batch_device(gray.data_ptr(), depth.data_ptr(), T.data_ptr(), F, 480, 640)
    st.sync()
    L.orbm_debug_read(out)
v = list(out)
n = max(v[10], 1)
print("CTA 0, cycles: stage+grid %d  projection %d  produce/resolve %d" % (v[0] // n, v[1] // n, v[2] // n))
print("producer warp 1: queries %d  wait-slot %d  walk+publish %d  (per query %d)  loop total %d" %
      (v[6], v[4], v[5], v[5] // max(v[6], 1), v[7]))
print("consumer: spin-wait %d  loop total %d" % (v[8], v[9]))
