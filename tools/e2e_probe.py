"""Host-buffer (end-to-end) throughput of the batch tracker under a few configurations (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from orb_slam2_ssd_semantic_b200 import PointCloudMapping, StreamTracker, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
KF = 12
ws = synth.WallStream(seed=1234, n=F)
fr = [ws.frame(t) for t in range(F)]
gray = torch.from_numpy(np.stack([f[0] for f in fr])).pin_memory()
depth = np.stack([f[1] for f in fr])
d16 = torch.from_numpy(np.rint(depth.astype(np.float64) * synth.DEPTH_FACTOR).astype(np.uint16)).pin_memory()
rgb = np.stack([f[2] for f in fr])
T = torch.from_numpy(np.ascontiguousarray(np.stack([f[3] for f in fr]), np.float32)).pin_memory()
kfs = list(range(0, F, KF))
kf_d16 = torch.from_numpy(np.ascontiguousarray(d16.numpy()[kfs])).pin_memory()
kf_rgb = torch.from_numpy(np.ascontiguousarray(rgb[kfs])).pin_memory()
factor = np.float32(1.0 / synth.DEPTH_FACTOR)
mk = lambda: StreamTracker(1000, 1.2, 8, 20, 7, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, max_frames=F)
trk = [mk(), mk(), mk()]
outs = [t.alloc_outputs(F, pinned=True) for t in trk]
pcm = PointCloudMapping(0.05)
Tk = T.numpy()[kfs]


def run(n, inflight, mapper, chunk, full_depth, chain=False):
    for t in trk:
        t.set_chunk_frames(chunk)
        t.set_full_depth_upload(full_depth)

    def submit(k):
        if mapper:
            pcm.insert_keyframes_u16(kf_d16.numpy(), kf_rgb.numpy(), factor, Tk, synth.FX, synth.FY, synth.CX, synth.CY)
        if chain and inflight > 1:
            trk[k % inflight].chain_after(trk[(k - 1) % inflight])
        trk[k % inflight].submit_batch_u16(gray.numpy(), d16.numpy(), factor, T.numpy(), outs[k % inflight])

    def go(m):
        for k in range(m):
            if k >= inflight:
                trk[k % inflight].sync()
            submit(k)
        for t in trk:
            t.sync()
        pcm.sync()

    go(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    go(n)
    dt = time.perf_counter() - t0
    return F * n / dt, dt / n * 1e3


cfgs = [(2, True, 128, False, False), (2, True, 128, False, True), (2, True, 256, False, True), (3, True, 256, False, True), (2, False, 256, False, True)] if len(sys.argv) > 2 else \
    [(i, m, c, f, False) for i in (1, 2) for m in (False, True) for c in (64, 128, 256) for f in (False, True)]
for inflight, mapper, chunk, full, chain in cfgs:
    trk[0].profile_enable(True)
    trk[0].profile_read()
    fps, ms = run(12, inflight, mapper, chunk, full, chain)
    print("inflight %d mapper %d chunk %3d full_depth %d chain %d : %8.0f frames/s  %.2f ms/batch" % (inflight, mapper, chunk, full, chain, fps, ms))
    st_ms, frames, runs = trk[0].profile_read()
    trk[0].profile_enable(False)
    print("   handle 0 stage ms per batch:", {k: round(v / max(runs, 1) * (F / max(frames / max(runs, 1), 1)), 3) for k, v in st_ms.items()}, "sum %.3f" % (sum(st_ms.values()) / max(frames, 1) * F))
