"""Batched stream pipeline (north_star's many-frame mode) over orbs_* of the C-ABI: per frame of a batch,
extract + ComputeStereoFromRGBD + UnprojectStereo of the previous frame + SearchByProjection(cur, last)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._abi import OrbsParams, OrbxParams, ptr
from .extractor import KP_DTYPE


class StreamTracker:
    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, fx=535.4, fy=539.2,
                 cx=320.1, cy=247.6, bf=40.0, th=15.0, nnratio=0.9, checkOri=True, max_frames=256, device=0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        p = OrbsParams(OrbxParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST), fx, fy, cx, cy, bf, th,
                       nnratio, int(checkOri), max_frames)
        _lib.check(self._L.orbs_create(C.byref(p), int(device), C.byref(self._h)))
        self.cap = int(self._L.orbx_max_keypoints(self._L.orbs_extractor(self._h)))
        self.nlevels = nlevels

    def __del__(self):
        try:
            if self._h:
                self._L.orbs_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def alloc_outputs(self, nframes: int, pinned: bool = False):
        """Output buffers for track_batch (kps, desc, nkp, cur2last, nmatch); pinned=True uses page-locked host
        memory (through torch) so the device->host copies are plain DMA."""
        cap = self.cap
        shapes = [((nframes, cap), KP_DTYPE), ((nframes, cap, 32), np.uint8), ((nframes,), np.int32),
                  ((nframes, cap), np.int32), ((nframes,), np.int32)]
        if not pinned:
            return tuple(np.zeros(s, d) for s, d in shapes)
        import torch
        outs = []
        self._pinned_keep = []
        for s, d in shapes:
            nbytes = int(np.prod(s)) * np.dtype(d).itemsize
            t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            self._pinned_keep.append(t)
            outs.append(t.numpy().view(d).reshape(s))
        return tuple(outs)

    def track_batch(self, gray: np.ndarray, depth: np.ndarray, Tcw: np.ndarray, out=None):
        """Host buffers in, host buffers out (the reference-facing call; copies are inside)."""
        gray = np.ascontiguousarray(gray, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(-1, 16)
        F, rows, cols = gray.shape
        assert depth.shape == gray.shape and Tcw.shape[0] == F
        cap = self.cap
        kps, desc, nkp, c2l, nm = out if out is not None else self.alloc_outputs(F)
        _lib.check(self._L.orbs_track_batch(self._h, ptr(gray), ptr(depth), ptr(Tcw), F, rows, cols, ptr(kps), ptr(desc),
                                            ptr(nkp), ptr(c2l), ptr(nm), cap))
        return kps, desc, nkp, c2l, nm

    def track_batch_u16(self, gray: np.ndarray, depth_u16: np.ndarray, depth_factor: float, Tcw: np.ndarray, out=None):
        """Like track_batch with the sensor's CV_16U depth; convertTo(CV_32F, depth_factor) runs on the device."""
        gray = np.ascontiguousarray(gray, np.uint8)
        depth_u16 = np.ascontiguousarray(depth_u16, np.uint16)
        Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(-1, 16)
        F, rows, cols = gray.shape
        assert depth_u16.shape == gray.shape and Tcw.shape[0] == F
        kps, desc, nkp, c2l, nm = out if out is not None else self.alloc_outputs(F)
        _lib.check(self._L.orbs_track_batch_u16(self._h, ptr(gray), ptr(depth_u16), float(np.float32(depth_factor)), ptr(Tcw),
                                                F, rows, cols, ptr(kps), ptr(desc), ptr(nkp), ptr(c2l), ptr(nm), self.cap))
        return kps, desc, nkp, c2l, nm

    def submit_batch_u16(self, gray: np.ndarray, depth_u16: np.ndarray, depth_factor: float, Tcw: np.ndarray, out):
        """Asynchronous track_batch_u16: returns after enqueueing; `out` (alloc_outputs(pinned=True)) is valid after
        sync().  The input arrays must be page-locked, C-contiguous and stay untouched until then."""
        for a, dt in ((gray, np.uint8), (depth_u16, np.uint16), (Tcw, np.float32)):
            if a.dtype != dt or not a.flags.c_contiguous:
                raise ValueError("submit_batch_u16 takes C-contiguous %s arrays (no hidden copies)" % np.dtype(dt).name)
        F, rows, cols = gray.shape
        assert depth_u16.shape == gray.shape and Tcw.size == F * 16
        kps, desc, nkp, c2l, nm = out
        _lib.check(self._L.orbs_submit_batch_u16(self._h, ptr(gray), ptr(depth_u16), float(np.float32(depth_factor)), ptr(Tcw),
                                                 F, rows, cols, ptr(kps), ptr(desc), ptr(nkp), ptr(c2l), ptr(nm), self.cap))
        return out

    def set_full_depth_upload(self, on: bool):
        """False (default): page-locked CV_16U depth is read under the keypoints in place; True: always upload it."""
        _lib.check(self._L.orbs_set_full_depth_upload(self._h, int(on)))

    def set_chunk_frames(self, frames: int):
        """Frames per upload chunk of the host-buffer calls (default 128)."""
        _lib.check(self._L.orbs_set_chunk_frames(self._h, int(frames)))

    def chain_after(self, prev: "StreamTracker"):
        """The next submitted batch starts its kernels when `prev`'s kernels are done (uploads are not held back)."""
        _lib.check(self._L.orbs_chain_after(self._h, prev._h))

    def device_inputs(self):
        """(d_gray, d_depth) device pointers of the last host-buffer batch (depth as f32 metres)."""
        g, d = C.c_void_p(), C.c_void_p()
        _lib.check(self._L.orbs_device_inputs(self._h, C.byref(g), C.byref(d)))
        return g.value, d.value

    def track_batch_device(self, d_gray: int, d_depth: int, d_Tcw: int, nframes: int, rows: int, cols: int):
        """Device pointers (ints) in; results stay in HBM (device_results()). Asynchronous."""
        _lib.check(self._L.orbs_track_batch_device(self._h, C.c_void_p(d_gray), C.c_void_p(d_depth), C.c_void_p(d_Tcw),
                                                   nframes, rows, cols))

    def device_results(self):
        ps = [C.c_void_p() for _ in range(5)]
        cap = C.c_int()
        _lib.check(self._L.orbs_device_results(self._h, *[C.byref(p) for p in ps], C.byref(cap)))
        return [p.value for p in ps], cap.value

    def frame_glue(self, frame: int, n: int):
        """(uright, depth, xw, valid) of the first n keypoints of `frame` of the last batch."""
        cap = self.cap
        ur, dp, xw, va = np.zeros(cap, np.float32), np.zeros(cap, np.float32), np.zeros((cap, 3), np.float32), np.zeros(cap, np.uint8)
        _lib.check(self._L.orbs_read_frame_glue(self._h, int(frame), ptr(ur), ptr(dp), ptr(xw), ptr(va), cap))
        return ur[:n], dp[:n], xw[:n], va[:n]

    STAGES = ("resize", "fast", "quadtree", "blur", "orient_desc", "glue", "match")

    def profile_enable(self, on: bool = True):
        _lib.check(self._L.orbx_profile_enable(self._L.orbs_extractor(self._h), int(on)))

    def profile_read(self):
        """-> ({stage: total ms}, frames, runs) since the last read (synchronises)."""
        ms = np.zeros(7, np.float32)
        fr, runs = C.c_longlong(0), C.c_int(0)
        _lib.check(self._L.orbx_profile_read(self._L.orbs_extractor(self._h), ptr(ms), C.byref(fr), C.byref(runs)))
        return dict(zip(self.STAGES, ms.tolist())), fr.value, runs.value

    def sync(self):
        _lib.check(self._L.orbs_sync(self._h))

    def stream(self) -> int:
        return int(self._L.orbs_stream(self._h) or 0)

    def launch_count(self) -> int:
        return int(self._L.orbs_launch_count(self._h))
