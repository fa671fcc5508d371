"""oracle/orb_cv2.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Line-by-line Python restatement of the reference extractor (src/ORBextractor.cc) that calls the REAL
OpenCV primitives through cv2 (4.13.0 in the build container) for everything the reference delegates to
OpenCV: cv2.resize (:1134), cv2.copyMakeBorder (:1136,1141), cv2.FAST (:818,823), cv2.GaussianBlur
(:1095), cv2.fastAtan2 (:87).  It exists to PIN the dependency-free C++ oracle (oracle/orb_ref.cpp) and to
generate the golden vectors under tests/golden/ (tools/make_golden.py).  Slow (pure-Python quad-tree
and descriptor loops): use on single frames only.

float32 discipline: every arithmetic step the reference does in `float` is done in np.float32 here.
"""
from __future__ import annotations

import math
import os

import numpy as np

import ctypes as _C

_LIBM = _C.CDLL("libm.so.6")
_LIBM.cosf.restype = _C.c_float
_LIBM.cosf.argtypes = [_C.c_float]
_LIBM.sinf.restype = _C.c_float
_LIBM.sinf.argtypes = [_C.c_float]

PATCH_SIZE = 31
HALF_PATCH_SIZE = 15
EDGE_THRESHOLD = 19
f32 = np.float32


def _pattern() -> np.ndarray:
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "orb_pattern_31.inc")
    txt = open(inc).read()
    txt = txt[txt.index("*/") + 2:]
    vals = [int(v) for v in txt.replace("\n", "").split(",") if v.strip()]
    assert len(vals) == 1024
    return np.array(vals, dtype=np.int32).reshape(512, 2)


def cv_round(v) -> int:
    """cvRound: round-half-to-even."""
    return int(np.rint(v))


class _Node:
    __slots__ = ("UL", "UR", "BL", "BR", "keys", "noMore", "seq", "alive")

    def __init__(self):
        self.keys = []
        self.noMore = False
        self.seq = 0
        self.alive = True


class ORBextractorCV2:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:41-118)."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        import cv2  # noqa: F401  (fail early if missing)
        self.nfeatures = int(nfeatures)
        self.scaleFactor = float(f32(scaleFactor))  # ctor arg is float, member double
        self.nlevels = int(nlevels)
        self.iniThFAST = int(iniThFAST)
        self.minThFAST = int(minThFAST)
        sf = [f32(1.0)]
        for i in range(1, nlevels):
            sf.append(f32(float(sf[i - 1]) * self.scaleFactor))
        self.mvScaleFactor = sf
        self.mvLevelSigma2 = [f32(s * s) for s in sf]
        self.mvInvScaleFactor = [f32(f32(1.0) / s) for s in sf]
        self.mvInvLevelSigma2 = [f32(f32(1.0) / s) for s in self.mvLevelSigma2]
        factor = f32(1.0 / self.scaleFactor)
        nd = f32(f32(f32(self.nfeatures) * f32(f32(1) - factor)) /
                 f32(f32(1) - f32(math.pow(float(factor), float(nlevels)))))
        self.mnFeaturesPerLevel = []
        s = 0
        for _ in range(nlevels - 1):
            n = cv_round(nd)
            self.mnFeaturesPerLevel.append(n)
            s += n
            nd = f32(nd * factor)
        self.mnFeaturesPerLevel.append(max(self.nfeatures - s, 0))
        self.pattern = _pattern()
        umax = [0] * (HALF_PATCH_SIZE + 1)
        vmax = int(math.floor(float(f32(f32(HALF_PATCH_SIZE) * f32(math.sqrt(2.0))) / f32(2) + f32(1))))
        vmin = int(math.ceil(float(f32(f32(HALF_PATCH_SIZE) * f32(math.sqrt(2.0))) / f32(2))))
        hp2 = float(HALF_PATCH_SIZE * HALF_PATCH_SIZE)
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(hp2 - v * v))
        v0 = 0
        for v in range(HALF_PATCH_SIZE, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax
        self.mvImagePyramid = [None] * nlevels
        self._bordered = [None] * nlevels
        self._seq = 0
        self.candidates_per_level = [0] * nlevels

    # -- ComputePyramid :1117-1145 ------------------------------------------------------------
    def ComputePyramid(self, image: np.ndarray):
        import cv2
        B = EDGE_THRESHOLD
        for level in range(self.nlevels):
            scale = self.mvInvScaleFactor[level]
            w = cv_round(f32(f32(image.shape[1]) * scale))
            h = cv_round(f32(f32(image.shape[0]) * scale))
            if level != 0:
                lvl = cv2.resize(self.mvImagePyramid[level - 1], (w, h), interpolation=cv2.INTER_LINEAR)
                temp = cv2.copyMakeBorder(lvl, B, B, B, B, cv2.BORDER_REFLECT_101 | cv2.BORDER_ISOLATED)
            else:
                temp = cv2.copyMakeBorder(image, B, B, B, B, cv2.BORDER_REFLECT_101)
            self._bordered[level] = temp
            self.mvImagePyramid[level] = temp[B:B + h, B:B + w]

    # -- DivideNode :478-534 --------------------------------------------------------------------
    @staticmethod
    def _divide(p: _Node):
        halfX = int(math.ceil(float(f32(p.UR[0] - p.UL[0]) / f32(2))))
        halfY = int(math.ceil(float(f32(p.BR[1] - p.UL[1]) / f32(2))))
        n1, n2, n3, n4 = _Node(), _Node(), _Node(), _Node()
        n1.UL = p.UL
        n1.UR = (p.UL[0] + halfX, p.UL[1])
        n1.BL = (p.UL[0], p.UL[1] + halfY)
        n1.BR = (p.UL[0] + halfX, p.UL[1] + halfY)
        n2.UL = n1.UR
        n2.UR = p.UR
        n2.BL = n1.BR
        n2.BR = (p.UR[0], p.UL[1] + halfY)
        n3.UL = n1.BL
        n3.UR = n1.BR
        n3.BL = p.BL
        n3.BR = (n1.BR[0], p.BL[1])
        n4.UL = n3.UR
        n4.UR = n2.BR
        n4.BL = n3.BR
        n4.BR = p.BR
        for kp in p.keys:
            if kp[0] < n1.UR[0]:
                (n1 if kp[1] < n1.BR[1] else n3).keys.append(kp)
            elif kp[1] < n1.BR[1]:
                n2.keys.append(kp)
            else:
                n4.keys.append(kp)
        for n in (n1, n2, n3, n4):
            if len(n.keys) == 1:
                n.noMore = True
        return n1, n2, n3, n4

    # -- DistributeOctTree :540-765 -------------------------------------------------------------
    # The std::list is modelled as a Python list in FRONT->BACK order where push_front = insert(0).
    def DistributeOctTree(self, keys, minX, maxX, minY, maxY, N):
        nIni = int(np.round(f32(maxX - minX) / f32(maxY - minY)))  # C round(): half away from zero; values >0
        nIni = int(math.floor(float(f32(maxX - minX) / f32(maxY - minY)) + 0.5))
        hX = f32(f32(maxX - minX) / f32(nIni))
        nodes = []
        ini = []
        for i in range(nIni):
            ni = _Node()
            ni.UL = (int(f32(hX * f32(i))), 0)
            ni.UR = (int(f32(hX * f32(i + 1))), 0)
            ni.BL = (ni.UL[0], maxY - minY)
            ni.BR = (ni.UR[0], maxY - minY)
            ni.seq = self._seq
            self._seq += 1
            nodes.append(ni)
            ini.append(ni)
        for kp in keys:
            ini[int(f32(kp[0]) / hX)].keys.append(kp)
        kept = []
        for n in nodes:
            if len(n.keys) == 1:
                n.noMore = True
                kept.append(n)
            elif len(n.keys) == 0:
                pass
            else:
                kept.append(n)
        nodes = kept
        finish = False
        vss = []

        def push_children(children, vss_out, count):
            for c in children:
                if len(c.keys) > 0:
                    c.seq = self._seq
                    self._seq += 1
                    nodes.insert(0, c)
                    if len(c.keys) > 1:
                        count[0] += 1
                        vss_out.append((len(c.keys), c.seq, c))

        while not finish:
            prevSize = len(nodes)
            nToExpand = [0]
            vss = []
            # sweep: visit the nodes that were in the list at sweep start, in list order
            snapshot = list(nodes)
            for nd in snapshot:
                if nd.noMore:
                    continue
                ch = self._divide(nd)
                push_children(ch, vss, nToExpand)
                nodes.remove(nd)
            if len(nodes) >= N or len(nodes) == prevSize:
                finish = True
            elif len(nodes) + nToExpand[0] * 3 > N:
                while not finish:
                    prevSize = len(nodes)
                    prev = sorted(vss, key=lambda t: (t[0], t[1]))
                    vss = []
                    for j in range(len(prev) - 1, -1, -1):
                        nd = prev[j][2]
                        ch = self._divide(nd)
                        push_children(ch, vss, [0])
                        nodes.remove(nd)
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prevSize:
                        finish = True
        out = []
        for n in nodes:
            best = n.keys[0]
            maxR = best[2]
            for k in n.keys[1:]:
                if k[2] > maxR:
                    best = k
                    maxR = k[2]
            out.append(best)
        return out

    # -- ComputeKeyPointsOctTree :771-862 -------------------------------------------------------
    def ComputeKeyPointsOctTree(self):
        import cv2
        allkps = []
        W = f32(30)
        for level in range(self.nlevels):
            img = self.mvImagePyramid[level]
            minBX = EDGE_THRESHOLD - 3
            minBY = minBX
            maxBX = img.shape[1] - EDGE_THRESHOLD + 3
            maxBY = img.shape[0] - EDGE_THRESHOLD + 3
            width = f32(maxBX - minBX)
            height = f32(maxBY - minBY)
            nCols = int(width / W)
            nRows = int(height / W)
            wCell = int(math.ceil(float(f32(width / f32(nCols)))))
            hCell = int(math.ceil(float(f32(height / f32(nRows)))))
            todist = []
            for i in range(nRows):
                iniY = f32(minBY + i * hCell)
                maxY = f32(iniY + f32(hCell + 6))
                if iniY >= maxBY - 3:
                    continue
                if maxY > maxBY:
                    maxY = f32(maxBY)
                for j in range(nCols):
                    iniX = f32(minBX + j * wCell)
                    maxX = f32(iniX + f32(wCell + 6))
                    if iniX >= maxBX - 6:
                        continue
                    if maxX > maxBX:
                        maxX = f32(maxBX)
                    roi = img[int(iniY):int(maxY), int(iniX):int(maxX)]
                    det = cv2.FastFeatureDetector_create(threshold=self.iniThFAST, nonmaxSuppression=True)
                    kps = det.detect(roi)
                    if len(kps) == 0:
                        det = cv2.FastFeatureDetector_create(threshold=self.minThFAST, nonmaxSuppression=True)
                        kps = det.detect(roi)
                    for kp in kps:
                        todist.append((f32(kp.pt[0]) + f32(j * wCell), f32(kp.pt[1]) + f32(i * hCell),
                                       f32(kp.response)))
            self.candidates_per_level[level] = len(todist)
            if todist:
                res = self.DistributeOctTree(todist, minBX, maxBX, minBY, maxBY, self.mnFeaturesPerLevel[level])
            else:
                res = []
            scaledPatchSize = int(f32(PATCH_SIZE) * self.mvScaleFactor[level])
            lv = []
            for (x, y, r) in res:
                lv.append([f32(x + f32(minBX)), f32(y + f32(minBY)), f32(scaledPatchSize), f32(-1), f32(r), level])
            allkps.append(lv)
        for level in range(self.nlevels):
            for kp in allkps[level]:
                kp[3] = self.IC_Angle(level, kp[0], kp[1])
        return allkps

    # -- IC_Angle :59-88 -------------------------------------------------------------------------
    def IC_Angle(self, level, x, y):
        import cv2
        B = EDGE_THRESHOLD
        buf = self._bordered[level].astype(np.int64)
        cy = cv_round(y) + B
        cx = cv_round(x) + B
        m_01 = 0
        m_10 = 0
        for u in range(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1):
            m_10 += u * int(buf[cy, cx + u])
        for v in range(1, HALF_PATCH_SIZE + 1):
            d = self.umax[v]
            us = np.arange(-d, d + 1)
            vp = buf[cy + v, cx - d:cx + d + 1]
            vm = buf[cy - v, cx - d:cx + d + 1]
            m_01 += v * int((vp - vm).sum())
            m_10 += int((us * (vp + vm)).sum())
        return f32(cv2.fastAtan2(float(f32(m_01)), float(f32(m_10))))

    # -- computeOrbDescriptor :92-131 -------------------------------------------------------------
    def computeOrbDescriptor(self, kp, img):
        factorPI = f32(math.pi / float(f32(180.0)))
        angle = f32(kp[3] * factorPI)
        # cos()/sin() on a float bind to libm's cosf/sinf (not correctly rounded): call the host libm itself
        a = f32(_LIBM.cosf(float(angle)))
        b = f32(_LIBM.sinf(float(angle)))
        cy = cv_round(kp[1])
        cx = cv_round(kp[0])
        px = self.pattern[:, 0].astype(np.float32)
        py = self.pattern[:, 1].astype(np.float32)
        yy = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int64)
        xx = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int64)
        vals = img[cy + yy, cx + xx].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)

    # -- operator() :1052-1114 ----------------------------------------------------------------------
    def __call__(self, image: np.ndarray):
        import cv2
        if image is None or image.size == 0:
            return np.zeros((0, 7), np.float32), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2
        self.ComputePyramid(np.ascontiguousarray(image))
        allkps = self.ComputeKeyPointsOctTree()
        kps_out = []
        desc_out = []
        for level in range(self.nlevels):
            kps = allkps[level]
            if not kps:
                continue
            work = self.mvImagePyramid[level].copy()
            work = cv2.GaussianBlur(work, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            for kp in kps:
                desc_out.append(self.computeOrbDescriptor(kp, work))
                x, y = kp[0], kp[1]
                if level != 0:
                    s = self.mvScaleFactor[level]
                    x = f32(x * s)
                    y = f32(y * s)
                kps_out.append((x, y, kp[2], kp[3], kp[4], kp[5], -1))
        n = len(kps_out)
        K = np.zeros(n, dtype=KP_DTYPE)
        for i, k in enumerate(kps_out):
            K[i] = k
        D = np.stack(desc_out).astype(np.uint8) if n else np.zeros((0, 32), np.uint8)
        return K, D


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
