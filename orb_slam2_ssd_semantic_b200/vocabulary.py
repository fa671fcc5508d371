"""Host-side mirror of ORB_SLAM2::ORBVocabulary (include/ORBVocabulary.h: DBoW2 TemplatedVocabulary<FORB>) for the one
call the hot path makes: transform(features, BowVector, FeatureVector, levelsup) from Frame::ComputeBoW /
KeyFrame::ComputeBoW (src/Frame.cc:546-555, src/KeyFrame.cc:75-84).  The tree walk runs on the GPU (orbv_transform);
the two maps are assembled here in feature order like TemplatedVocabulary::transform does (TF-IDF weights, L1 norm)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._abi import ptr


class ORBVocabulary:
    def __init__(self, k: int, L: int, parent, node_desc, weight, word_id=None, device: int = 0):
        """node 0 = root, parent[i] < i, children in ascending id; leaves are the words (ascending id unless word_id)."""
        self._L = _lib.lib()
        self._h = C.c_void_p()
        parent = np.ascontiguousarray(parent, np.int32)
        node_desc = np.ascontiguousarray(node_desc, np.uint8).reshape(-1, 32)
        weight = np.ascontiguousarray(weight, np.float64)
        wid = None if word_id is None else np.ascontiguousarray(word_id, np.int32)
        assert len(parent) == len(node_desc) == len(weight)
        _lib.check(self._L.orbv_create(int(device), int(k), int(L), len(parent), ptr(parent), ptr(node_desc), ptr(weight),
                                       ptr(wid), C.byref(self._h)))
        self.k, self.L = k, L

    def __del__(self):
        try:
            if self._h:
                self._L.orbv_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def transform_raw(self, desc, levelsup: int = 4):
        """Per feature: (word id, word weight, node id `levelsup` levels above the leaves)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word = np.zeros(max(n, 1), np.uint32)
        w = np.zeros(max(n, 1), np.float64)
        node = np.zeros(max(n, 1), np.uint32)
        _lib.check(self._L.orbv_transform(self._h, ptr(desc), n, int(levelsup), ptr(word), ptr(w), ptr(node)))
        return word[:n], w[:n], node[:n]

    def transform(self, desc, levelsup: int = 4):
        """-> (BowVector: dict word -> value, L1-normalised; FeatureVector: dict node -> [feature indices])."""
        word, w, node = self.transform_raw(desc, levelsup)
        bow, fv = {}, {}
        for i in range(len(word)):
            if not w[i] > 0:
                continue
            wi = int(word[i])
            bow[wi] = bow.get(wi, 0.0) + float(w[i])          # BowVector::addWeight: running double sum in feature order
            fv.setdefault(int(node[i]), []).append(i)
        norm = 0.0
        for k in sorted(bow):                                   # BowVector::normalize(L1): map order
            norm += abs(bow[k])
        if norm > 0.0:
            for k in bow:
                bow[k] /= norm
        return bow, fv

    def num_words(self) -> int:
        return int(self._L.orbv_num_words(self._h))
