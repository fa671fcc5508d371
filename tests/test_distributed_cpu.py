"""CPU suite for the N>1 logic (gloo, world_size 2): frame sharding, the variable-length summary all-gather, and
the clamp-add composition algebra that makes the sharded occupancy map equal the sequential one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orb_slam2_ssd_semantic_b200 import distributed as D


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                b, e = D.shard_range(n, world, r)
                assert 0 <= b <= e <= n
                got.extend(range(b, e))
                hb, b2, e2 = D.shard_frames_with_halo(n, world, r)
                assert (b2, e2) == (b, e) and hb == max(b - 1, 0)
            assert got == list(range(n))


def _seq_apply(v, deltas, cmin, cmax):
    for d in deltas:
        v = np.float32(min(max(np.float32(v + d), cmin), cmax))
    return v


def test_clamp_add_composition_equals_sequential(oracle):
    hit, miss, cmin, cmax = [np.float32(v) for v in oracle.RefOccupancy().constants()]
    rng = np.random.default_rng(0)
    for _ in range(300):
        seqs = [rng.choice([hit, miss], size=rng.integers(0, 12)) for _ in range(3)]   # three shards of updates
        # per-shard summary, built exactly like map_update does (ocm.cu)
        summ = []
        for s in seqs:
            a, lo, hi = np.float32(0), np.float32(-np.inf), np.float32(np.inf)
            for d in s:
                a = np.float32(a + d)
                lo = np.float32(min(max(np.float32(lo + d), cmin), cmax))
                hi = np.float32(min(max(np.float32(hi + d), cmin), cmax))
            summ.append((a, lo, hi))
        v = np.float32(0)
        for (a, lo, hi) in summ:    # apply shards in order (k_ocm_apply_summaries)
            v = np.float32(min(max(np.float32(v + a), lo), hi))
        ref = _seq_apply(np.float32(0), np.concatenate(seqs) if sum(map(len, seqs)) else [], cmin, cmax)
        assert abs(float(v) - float(ref)) <= 1e-5
        # and the closed form of composing two summaries
        a12, lo12, hi12 = D.compose_summaries(*summ[0], *summ[1])
        x = np.float32(rng.uniform(-2, 3))
        two = min(max(min(max(x + summ[0][0], summ[0][1]), summ[0][2]) + summ[1][0], summ[1][1]), summ[1][2])
        one = min(max(x + a12, lo12), hi12)
        assert abs(float(two) - float(one)) <= 1e-5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n = 5 + 3 * rank     # different lengths per rank
        keys = torch.from_numpy(rng.integers(0, 2 ** 40, size=n).astype(np.int64))
        a = torch.from_numpy(rng.normal(size=n).astype(np.float32))
        lo = torch.from_numpy(rng.normal(size=n).astype(np.float32))
        hi = lo + 1
        shards = D.all_gather_summaries(keys, a, lo, hi)
        ok = len(shards) == world
        for r, (k2, a2, lo2, hi2) in enumerate(shards):
            rr = np.random.default_rng(100 + r)
            nn = 5 + 3 * r
            ok &= bool((k2.numpy() == rr.integers(0, 2 ** 40, size=nn).astype(np.int64)).all())
            ok &= bool(np.allclose(a2.numpy(), rr.normal(size=nn).astype(np.float32)))
        b, e = D.shard_range(257, world, rank)
        t = torch.tensor([e - b], dtype=torch.int64)
        dist.all_reduce(t)
        ok &= int(t.item()) == 257
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allgather_and_sharding():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bench_strong_scaling_partition():
    """bench.py's frame shards at 1 / 2 / 3 / 4 / 8 ranks: every frame of a step is owned by exactly one rank, pieces stay
    inside their batch, a tracked piece (owned frames + the halo frame) never exceeds a batch, keyframes are partitioned."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    total = bench.SUB * bench.BATCH
    for world in (1, 2, 3, 4, 8):
        owner = np.full(total, -1)
        nkf = 0
        for rank in range(world):
            pieces = bench.rank_pieces(world, rank)
            assert pieces, (world, rank)
            for (b, lo, hi) in pieces:
                assert 0 <= lo < hi <= bench.BATCH
                assert hi - max(lo - 1, 0) <= bench.BATCH
                assert (owner[b * bench.BATCH + lo:b * bench.BATCH + hi] == -1).all()
                owner[b * bench.BATCH + lo:b * bench.BATCH + hi] = rank
                nkf += sum(1 for t in range(lo, hi) if t % bench.KF_EVERY == 0)
            # contiguous global range
            g = [b * bench.BATCH + lo for (b, lo, hi) in pieces] + [pieces[-1][0] * bench.BATCH + pieces[-1][2]]
            assert all(x <= y for x, y in zip(g, g[1:]))
        assert (owner >= 0).all() and (np.diff(owner) >= 0).all()
        assert nkf == bench.SUB * len(range(0, bench.BATCH, bench.KF_EVERY))
