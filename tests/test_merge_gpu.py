"""GPU, 2 ranks over NCCL: keyframes sharded across GPUs, per-rank occupancy maps merged through the all-gather of
clamp-add summaries == the sequential single-map result of the CPU oracle (log-odds within 1e-5, same leaf set)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene(n):
    from orb_slam2_ssd_semantic_b200 import synth
    ws = synth.WallStream(seed=7, n=n, depth0=2.0)
    out = []
    for t in range(n):
        gray, depth, rgb, T = ws.frame(t * 9)
        yy, xx = np.mgrid[0:480, 0:640]
        depth = (depth + 0.25 * np.sin(xx / 80.0 + t) * np.cos(yy / 60.0)).astype(np.float32)
        out.append((depth, rgb, T))
    return out


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from orb_slam2_ssd_semantic_b200 import PointCloudMapping, synth
    from orb_slam2_ssd_semantic_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        scene = _scene(6)
        b, e = D.shard_range(len(scene), world, rank)
        pcm = PointCloudMapping(0.05, device=rank)
        label = np.zeros((480, 640), np.uint8)
        label[320:, :] = 1
        for depth, rgb, T in scene[b:e]:
            pcm.insertKeyFrame(T, depth, rgb, synth.FX, synth.FY, synth.CX, synth.CY, label)
        merged, nbytes, _ = D.merge_occupancy(pcm, device=torch.device("cuda", rank))
        keys, lo, _ = merged.export_leaves()
        q.put((rank, keys, lo, nbytes))
    finally:
        dist.destroy_process_group()


def test_two_gpu_map_merge_equals_sequential(oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from orb_slam2_ssd_semantic_b200 import synth
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
    ref = oracle.RefOccupancy()
    label = np.zeros((480, 640), np.uint8)
    label[320:, :] = 1
    for depth, rgb, T in _scene(6):
        ref.insert_keyframe(T, depth, rgb, synth.FX, synth.FY, synth.CX, synth.CY, label)
    kr, lr = ref.export_leaves()

    def pack(k, v):
        k = k.astype(np.uint64)
        p = k[:, 0] | (k[:, 1] << np.uint64(16)) | (k[:, 2] << np.uint64(32))
        o = np.argsort(p)
        return p[o], v[o]
    pr, vr = pack(kr, lr)
    for rank, keys, lo, nbytes in res:
        pg, vg = pack(keys, lo)
        assert len(pg) == len(pr) and (pg == pr).all(), "rank %d leaf set differs" % rank
        assert np.abs(vg - vr).max() <= 1e-5
        assert nbytes > 0
