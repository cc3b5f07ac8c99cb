"""Re-pin oracle/spec.py against the live reference (only where /root/reference is mounted, i.e.
in the build container; skipped on the GPU box).  CPU only."""
import numpy as np
import pytest

from oracle import ref_shim, spec

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")


def _to_arr(labels, ids):
    out = np.full(len(ids), -1, np.int64)
    pos = {int(i): k for k, i in enumerate(ids.tolist())}
    for i, c in labels.items():
        out[pos[int(i)]] = c
    return out


def test_live_battery():
    from oracle.make_golden import battery_case
    ns = ref_shim.load()
    rng = np.random.default_rng(777)
    for _ in range(40):
        mat, eps, mp = battery_case(rng)
        X, Y = mat[:, 1], mat[:, 2]
        assert np.array_equal(_to_arr(ns.cDBSCAN2(mat, eps, mp).labels, mat[:, 0]), spec.cdbscan_v2(X, Y, eps, mp))
        assert np.array_equal(_to_arr(ns.cDBSCAN(mat, eps, mp).labels, mat[:, 0]), spec.cdbscan_v1(X, Y, eps, mp))
        assert np.array_equal(_to_arr(ns.blockDBSCAN(mat, eps, mp).labels, mat[:, 0]), spec.blockdbscan(X, Y, eps, mp))
