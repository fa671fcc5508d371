"""Every kernel formulation / launch shape b200orb_set_tuning() can select yields the same bytes (GPU): the round-1 and
round-2 formulations of orientation + descriptor and of the FAST tile staging, FAST CTAs of 1 / 2 / 4 / 8 cells, the
quad-tree's three register budgets.  The process-wide setting is restored whatever happens."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _get(L):
    m, w, q = C.c_int(), C.c_int(), C.c_int()
    assert L.b200orb_get_tuning(C.byref(m), C.byref(w), C.byref(q)) == 0
    return m.value, w.value, q.value


def test_all_tunings_give_identical_results():
    from orb_slam2_ssd_semantic_b200 import ORBextractor, _lib, synth
    L = _lib.lib()
    saved = _get(L)
    imgs = np.stack([synth.synth_frame(1234, t) for t in range(3)])
    odd = np.ascontiguousarray(synth.synth_frame(77, 0)[:401, :533])        # odd geometry: other cell sizes / level shapes
    try:
        results = {}
        configs = [(0, 8, 2), (1, 8, 2), (2, 8, 2), (3, 8, 2), (3, 4, 2), (3, 2, 3), (3, 1, 4), (0, 4, 4), (3, 8, 4), (3, 8, 3)]
        for cfg in configs:
            _lib.check(L.b200orb_set_tuning(*cfg))
            assert _get(L) == cfg
            out = []
            for nfeat in (1000, 2000):
                ex = ORBextractor(nfeat, 1.2, 8, 20, 7)
                for k, d in ex.extract_batch(imgs):
                    out.append((k.tobytes(), d.tobytes()))
                k, d = ex(odd)
                out.append((k.tobytes(), d.tobytes()))
            results[cfg] = out
        base = results[configs[0]]
        assert all(len(k) > 0 for k, _ in base)
        for cfg in configs[1:]:
            assert results[cfg] == base, cfg
    finally:
        _lib.check(L.b200orb_set_tuning(*saved))
    assert _get(L) == saved


def test_bad_tuning_is_rejected():
    from orb_slam2_ssd_semantic_b200 import _lib
    L = _lib.lib()
    saved = _get(L)
    assert L.b200orb_set_tuning(-1, 3, -1) != 0 and L.b200orb_set_tuning(-1, -1, 5) != 0
    assert L.b200orb_set_tuning(-1, -1, -1) == 0
    assert _get(L) == saved
