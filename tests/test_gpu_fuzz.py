"""Differential fuzzing on the GPU: many small seeded inputs (ties, duplicates, tiny eps, negative
coordinates, cut filters) through the C ABI against oracle/spec.py, all three variants, labels and
candidate records.  Fixed seeds, so a failure names its case."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import spec  # noqa: E402

SPEC = {1: spec.cdbscan_v1, 2: spec.cdbscan_v2, 3: spec.blockdbscan}


def _case(rng):
    n = int(rng.integers(1, 900))
    span = int(rng.choice([40, 400, 5000, 60000, 2_000_000]))
    eps = int(rng.choice([1, 2, 3, 7, 30, 150, 900, 4000]))
    mp = int(rng.integers(1, 9))
    q = int(rng.choice([1, 1, 3, 10, 40]))
    mode = int(rng.integers(0, 4))
    if mode == 0:
        X = rng.integers(0, span, n)
        Y = X + rng.integers(0, span, n)
    elif mode == 1:                                           # tight clumps
        k = max(1, n // 25)
        cx, cy = rng.integers(0, span, k), rng.integers(0, span, k)
        w = rng.integers(0, k, n)
        X = cx[w] + rng.integers(-2 * eps, 2 * eps + 1, n)
        Y = cy[w] + rng.integers(-2 * eps, 2 * eps + 1, n)
    elif mode == 2:                                           # everything on few lattice sites
        X = rng.integers(0, 6, n) * eps
        Y = rng.integers(0, 6, n) * eps + rng.integers(0, 2, n)
    else:                                                     # negative and mixed signs
        X = rng.integers(-span, span, n)
        Y = rng.integers(-span, span, n)
    X, Y = (X // q) * q, (Y // q) * q
    cut = int(rng.choice([0, 0, eps, span // 3 + 1]))
    return X.astype(np.int64), Y.astype(np.int64), eps, mp, cut


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16])
def test_fuzz_labels_and_records(seed):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cloops_b200 import device
    rng = np.random.default_rng(seed)
    for case in range(150):
        X, Y, eps, mp, cut = _case(rng)
        dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
        act = (Y - X) >= cut if cut > 0 else np.ones(len(X), bool)
        for v in (2, 1, 3):
            want = np.full(len(X), -1, np.int64)
            if act.any():
                want[act] = SPEC[v](X[act], Y[act], eps, mp)
            got, info = device.dbscan_device(dx, dy, eps, mp, v, cut)
            got = got.cpu().numpy()
            assert np.array_equal(got, want), (seed, case, v, eps, mp, cut, len(X))
            p = device.Pass(dx, dy, eps, mp, v, cut, score=False)
            inter, selfl, in_i, in_s = spec.cluster_records(X, Y, want)
            bbox, kind = p.bbox.cpu().numpy(), p.kind.cpu().numpy()
            assert np.array_equal(bbox[kind == 1], inter[:, :4]) and np.array_equal(bbox[kind == 2], selfl[:, :4]), (seed, case, v)
            mk = p.member_kind.cpu().numpy()
            assert (int((mk == 1).sum()), int((mk == 2).sum())) == (int(in_i.sum()), int(in_s.sum())), (seed, case, v)
            p.close()
