#!/usr/bin/env python3
"""bench.py -- RGB-D frames/s of the ORB-SLAM2 hot path (extract + SearchByProjection match [+ octomap insert])
at 640x480 on N x B200, with the per-kernel HBM roofline and the CPU baseline beside it.

A "step" is one pass of the hot path over one batch of synthetic RGB-D frames (BASELINE.json configs[1]: a TUM
fr3_walking-shaped 640x480 stream, ORBextractor(1000,1.2,8,20,7), SearchByProjection th=15 against the previous
frame; every KF_EVERY-th frame is a keyframe pushed into the occupancy map).  One process per GPU; frames are
sharded across ranks with no data-path collective (weak scaling: every rank runs a full batch).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl b200|reference]

value   : whole-job frames/s with the batch already resident in HBM, timed with CUDA events on the pipeline's
          stream, max over ranks.
e2e     : the same metric through the reference-facing C-ABI call with HOST buffers (pinned), host<->device
          copies inside the timed region.
roofline: dominant kernel of the step, algorithmic bytes / CUDA-event duration measured live.
cpu_baseline / --impl reference: the CPU oracle (line-faithful port of the reference path; the reference itself
          cannot be compiled here) timed on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from orb_slam2_ssd_semantic_b200 import synth  # noqa: E402

ROWS, COLS = 480, 640
NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH = 1000, 1.2, 8, 20, 7    # perfect/Examples/RGB-D/TUM3.yaml:45-54
TH, NNRATIO = 15.0, 0.9                                        # src/Tracking.cc:1327,1346
KF_EVERY = 12   # tool/KeyFrameTrajectory_f3_walk_src.txt holds 69 keyframes for 827 frames -> 1 in 12
S_IN = ROWS * COLS
LEVEL_PX = [640 * 480, 533 * 400, 444 * 333, 370 * 278, 309 * 231, 257 * 193, 214 * 161, 179 * 134]
S_PYR = sum(LEVEL_PX)
METRIC = "RGB-D frames/sec (extract+match+octomap) @640x480"


def algorithmic_bytes(n_kp: float, n_cand: float) -> dict:
    """Per-frame algorithmic bytes of each stage."""
    return {
        "resize": (S_PYR - LEVEL_PX[-1]) + (S_PYR - LEVEL_PX[0]),
        "fast": S_PYR + 4 * n_cand,
        "quadtree": 4 * n_cand + 4 * n_kp,
        "blur": 2 * S_PYR,
        # per keypoint: the 749 raw pixels of the r = 15 disc (IC_Angle), the 512 blurred sample pixels of the pattern,
        # 4 B selection in, 60 B KeyPoint + descriptor out
        "orient_desc": (749 + 512 + 4 + 60) * n_kp,
        "glue": 69 * n_kp,
        "match": 52 * n_kp + 52 * n_kp + 4 * (64 * 48 + 1) + 8 * n_kp,
    }


def pipeline_bytes(n_kp: float) -> float:
    """B_ext + B_match of SURVEY §8(d)."""
    return S_IN + 5 * S_PYR + 60 * n_kp + (52 * n_kp + 52 * n_kp + 4 * (64 * 48 + 1) + 8 * n_kp)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_batch(nframes: int, seed: int = 1234):
    ws = synth.WallStream(seed=seed, n=nframes)
    gray = np.empty((nframes, ROWS, COLS), np.uint8)
    depth = np.empty((nframes, ROWS, COLS), np.float32)
    rgb = np.empty((nframes, ROWS, COLS, 3), np.uint8)
    T = np.empty((nframes, 4, 4), np.float32)
    for t in range(nframes):
        gray[t], depth[t], rgb[t], T[t] = ws.frame(t)
    return gray, depth, rgb, T


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of src/ORBextractor.cc + src/Frame.cc glue + src/ORBmatcher.cc) on the host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_pipeline(gray, depth, rgb, T, nthreads: int):
    """extract -> stereo/unproject -> SearchByProjection(cur,last) over the frames on `nthreads` host threads
    (oracle/pipeline_ref.cpp: frame-parallel, one extractor instance per thread as src/Frame.cc:121-124 does for
    stereo) with every KF_EVERY-th frame also pushed through the occupancy oracle on its own mapping thread (the
    reference maps on a separate std::thread, src/pointcloudmapping.cc:43); timed inside with steady_clock."""
    from oracle import ref
    sec, _, _ = ref.pipeline_run(gray, depth, T, nthreads, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, synth.FX, synth.FY,
                                 synth.CX, synth.CY, synth.BF, TH, NNRATIO, True, 1, rgb=rgb, kf_every=KF_EVERY)
    return sec


def run_reference(args):
    """--impl reference: the CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    nthreads = min(cores, 64)
    sample = min(args.frames, max(2 * nthreads, 32))
    gray, depth, rgb, T = make_batch(sample)
    for _ in range(min(args.warmup, 1)):
        cpu_pipeline(gray[:max(2, nthreads)], depth, rgb, T, nthreads)
    times = [cpu_pipeline(gray, depth, rgb, T, nthreads) for _ in range(args.steps)]
    tot = float(np.sum(times))
    value = sample * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, sample),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": nthreads, "kind": "port",
                         "sample": "%d frames/step x %d steps, frame-parallel on %d threads" % (sample, args.steps, nthreads)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# DRAM traffic of the single-launch stages, bytes per frame, from the ncu capture summarised in
# profiles/r01_ncu_v5_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum of a 256-frame launch / 256)
NCU_DRAM_BYTES_PER_FRAME = {"fast": (231405824 + 41911808) / 256.0, "quadtree": (74866176 + 93454336) / 256.0,
                            "orient_desc": (463207168 + 18829056) / 256.0}


def workload_config(args, frames):
    return {"workload": "TUM fr3_walking-shaped synthetic RGB-D stream 640x480, ORBextractor(1000,1.2,8,20,7) + "
                        "SearchByProjection(cur,last,th=15) per frame, every %dth frame a keyframe inserted into the "
                        "0.05 m occupancy map (BASELINE.json configs[1] + keyframe path of configs[3])" % KF_EVERY,
            "keyframes_per_step_per_gpu": len(range(0, frames, KF_EVERY)),
            "frames_per_step_per_gpu": frames, "nfeatures": NFEAT, "parallelism": "frame-sharded x%d" % args.gpus,
            "l2": ("inputs (%.0f MB gray+depth per step) exceed the 126 MB L2" if frames * S_IN * 5 > 126e6 else
                   "inputs (%.0f MB gray+depth per step) FIT the 126 MB L2: not a valid timing configuration, use the "
                   "default --frames") % (frames * S_IN * 5 / 1e6)}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from orb_slam2_ssd_semantic_b200 import PointCloudMapping, StreamTracker
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    F = args.frames
    gray, depth, rgb, T = make_batch(F, seed=1234 + rank)
    st = StreamTracker(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, TH,
                       NNRATIO, True, F, device=local)
    d_gray = torch.from_numpy(gray).to(dev)
    d_depth = torch.from_numpy(depth).to(dev)
    d_T = torch.from_numpy(T).to(dev)
    d_rgb = torch.from_numpy(rgb).to(dev)
    pcm = PointCloudMapping(0.05, device=local)
    kfs = list(range(0, F, KF_EVERY))
    ext = torch.cuda.ExternalStream(st.stream(), device=dev)
    ext_map = torch.cuda.ExternalStream(pcm.stream(), device=dev)
    npx = ROWS * COLS

    def step_device():
        # tracking (extract + glue + match) on the pipeline stream, dense mapping on its own stream beside it -- the
        # reference also maps on a separate thread (src/pointcloudmapping.cc:43)
        st.track_batch_device(d_gray.data_ptr(), d_depth.data_ptr(), d_T.data_ptr(), F, ROWS, COLS)
        pcm.insert_keyframes_device(d_depth.data_ptr(), d_rgb.data_ptr(), ROWS, COLS, kfs, T[kfs], synth.FX, synth.FY,
                                    synth.CX, synth.CY)

    def join_streams():
        ev = torch.cuda.Event()
        ev.record(ext_map)
        ext.wait_event(ev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    st.sync()
    pcm.sync()
    launches0 = st.launch_count() + pcm.launch_count()
    st.profile_enable(True)
    st.profile_read()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    em0, em1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    em0.record(ext_map)
    for _ in range(args.steps):
        step_device()
    em1.record(ext_map)
    join_streams()
    e1.record(ext)
    st.sync()
    pcm.sync()
    barrier()
    map_ms = em0.elapsed_time(em1)
    clocks = sampler.result()
    ms_total = e0.elapsed_time(e1)
    stage_ms, prof_frames, prof_runs = st.profile_read()
    st.profile_enable(False)
    launches = st.launch_count() + pcm.launch_count() - launches0
    tms = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms.item())
    value = world * F * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers (pinned) through the reference-facing calls ----
    # What the host hands over per step is what the reference's callers hold: gray u8 + the sensor's CV_16U depth +
    # poses for every frame (Tracking::GrabImageRGBD converts the depth itself, src/Tracking.cc:366-367) and the
    # colour image of every keyframe (Tracking::CreateNewKeyFrame -> insertKeyFrame, src/Tracking.cc:1889).  The
    # tracker never uploads depth images (it reads the pixels under the keypoints in place); the mapper uploads the
    # depth + colour of the keyframes only.
    depth_u16 = np.rint(depth.astype(np.float64) * synth.DEPTH_FACTOR).astype(np.uint16)
    assert (depth_u16.astype(np.float32) * np.float32(1.0 / synth.DEPTH_FACTOR) == depth).all()
    p_gray = torch.from_numpy(gray).pin_memory()
    p_d16 = torch.from_numpy(depth_u16).pin_memory()
    p_T = torch.from_numpy(T).pin_memory()
    p_rgbk = torch.from_numpy(np.ascontiguousarray(rgb[kfs])).pin_memory()
    p_d16k = torch.from_numpy(np.ascontiguousarray(depth_u16[kfs])).pin_memory()
    kf_d16, kf_rgb = p_d16k.numpy(), p_rgbk.numpy()
    st_gray, st_d16, st_T = p_gray.numpy(), p_d16.numpy(), p_T.numpy()
    outs = st.alloc_outputs(F, pinned=True)
    factor = np.float32(1.0 / synth.DEPTH_FACTOR)
    # Two tracker handles used alternately keep two batches in flight: the upload of batch k+1 and the download of
    # batch k-1 overlap the kernels of batch k (orbs_chain_after orders the kernels of consecutive batches).  Every batch still crosses PCIe in both directions inside the timed
    # region, and its results are on the host (sync of its handle) before the batch after the next is submitted.
    st2 = StreamTracker(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, synth.FX, synth.FY, synth.CX, synth.CY, synth.BF, TH,
                        NNRATIO, True, F, device=local)
    trk = [st, st2]
    for t in trk:
        t.set_chunk_frames(F)   # whole-batch uploads: they overlap the other handle's kernels, not this handle's own
    outs2 = [outs, st2.alloc_outputs(F, pinned=True)]

    def submit(k):
        # mapper first, asynchronously on its own stream (ocm_insert_keyframes_u16): H2D depth (CV_16U) + colour of the
        # keyframes only, converted on the device, then the keyframe inserts
        pcm.insert_keyframes_u16(kf_d16, kf_rgb, factor, T[kfs], synth.FX, synth.FY, synth.CX, synth.CY)
        # tracker: H2D gray in chunks overlapped with extraction; the page-locked CV_16U depth is read under the
        # keypoints in place (zero-copy gather); D2H keypoints, descriptors, matches
        trk[k & 1].chain_after(trk[(k + 1) & 1])   # kernels of batch k start when those of batch k-1 are done
        trk[k & 1].submit_batch_u16(st_gray, st_d16, factor, st_T, outs2[k & 1])

    def collect(k):
        trk[k & 1].sync()
        return outs2[k & 1]

    def run_host(n):
        submit(0)
        for k in range(1, n):
            submit(k)
            collect(k - 1)
        o = collect(n - 1)
        pcm.sync()
        return o

    run_host(3)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(10, args.steps)   # enough batches that filling / draining the two-deep pipeline is amortised
    out = run_host(e2e_steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / float(te.item())
    kps, desc, nkp, c2l, nm = out
    # gray + poses + one 32-byte PCIe sector per keypoint of the in-place depth gather + keyframe depth (u16) and colour
    h2d = gray.nbytes + T.nbytes + int(nkp.sum()) * 32 + p_d16k.numel() * 2 + p_rgbk.numel()
    d2h = kps.nbytes + desc.nbytes + nkp.nbytes + c2l.nbytes + nm.nbytes
    n_kp = float(nkp.mean())
    n_match = float(nm[1:].mean())

    if rank == 0:
        from orb_slam2_ssd_semantic_b200 import _lib
        import ctypes as C
        cand = np.zeros(NLEVELS, np.int32)
        _lib.lib().orbx_candidates_per_level(_lib.lib().orbs_extractor(st._h), 0, cand.ctypes.data_as(C.c_void_p))
        n_cand = float(cand.sum())
        ab = algorithmic_bytes(n_kp, n_cand)
        peak, peak_src = measured_peak()
        stages = {}
        for k, v in stage_ms.items():
            per_launch_ms = v / max(prof_runs, 1)
            gbs = ab[k] * F / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
            stages[k] = {"ms_per_step": per_launch_ms, "algorithmic_bytes_per_frame": ab[k], "gbs": gbs,
                         "frac": gbs / peak}
        dom = max(stage_ms, key=lambda k: stage_ms[k])
        pipe_gbs = pipeline_bytes(n_kp) * F * args.steps / (ms_total * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, F),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "batches_in_flight": 2, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": stages[dom]["frac"],
                         "traffic": (NCU_DRAM_BYTES_PER_FRAME[dom] * F if dom in NCU_DRAM_BYTES_PER_FRAME else None),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture at 256 "
                                           "frames/launch (profiles/r01_ncu_v5_summary.txt), scaled to this launch",
                         "peak_source": peak_src,
                         "pipeline": {"achieved": pipe_gbs, "frac": pipe_gbs / peak,
                                      "algorithmic_bytes_per_frame": pipeline_bytes(n_kp)},
                         "stages": stages},
            "stats": {"keypoints_per_frame": n_kp, "fast_candidates_frame0": n_cand, "matches_per_frame": n_match,
                      "map_leaves": pcm.num_leaves(), "mapping_stream_ms_per_step": map_ms / args.steps,
                      "keyframes_per_step": len(kfs)},
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            nthreads = min(cores, 64)
            sample = min(F, max(2 * nthreads, 32))
            t_cpu = cpu_pipeline(gray[:sample], depth[:sample], rgb[:sample], T[:sample], nthreads)
            line["cpu_baseline"] = {"value": sample / t_cpu, "unit": "frames/s", "cores": nthreads, "kind": "port",
                                    "sample": "%d frames, frame-parallel on %d host threads (oracle/ C++ port)" % (sample, nthreads)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=256, help="frames per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
