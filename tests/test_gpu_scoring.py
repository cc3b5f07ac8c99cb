"""GPU parity of the permuted-background range counts (cModel.py:60-143): the 123 integers per
candidate against the reference's own sets (golden) and against oracle/spec.py."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import spec  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cloops_b200 import device
    return device


def test_range_counts_golden(dev, gold_dir):
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    pipe = np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))
    cov = dev.Coverage(dev.to_device_i32(d["X"]), dev.to_device_i32(d["Y"]))
    recs = pipe["sig_records"].copy()
    recs[:, 0] = np.maximum(recs[:, 0], 0)
    recs[:, 2] = np.maximum(recs[:, 2], 0)
    got = cov.range_counts(recs)
    want = pipe["sig_ints200"]
    assert np.array_equal(got[:200], want)
    # ra, rb, rab of every candidate against the tuples the reference returned
    tup = pipe["sig_tuples"]
    assert np.array_equal(got[:, :3], tup[:, 5:8].astype(np.int64))
    assert np.array_equal(cov.region_pets(recs), got[:, :3])


def test_range_counts_random(dev):
    rng = np.random.default_rng(5)
    n = 50000
    X = rng.integers(0, 200000, n)
    Y = X + rng.integers(0, 50000, n)
    X[:500] = rng.integers(0, 300, 500)          # PETs near 0: windows clamp at 0 (cModel.py:98-102)
    cov = dev.Coverage(dev.to_device_i32(X), dev.to_device_i32(Y))
    cands = []
    for _ in range(60):
        a0 = int(rng.integers(0, 150000)); a1 = a0 + int(rng.integers(0, 4000))
        b0 = a0 + int(rng.integers(0, 30000)); b1 = b0 + int(rng.integers(0, 4000))     # overlapping and nested hulls
        cands.append([a0, a1, b0, b1])
    cands += [[0, 50, 100, 400], [10, 10, 10, 10], [0, 0, 0, 0], [199000, 260000, 240000, 300000], [5, 2000, 3, 1500]]
    got = cov.range_counts(cands)
    for k, c in enumerate(cands):
        want = spec.range_counts(X, Y, c[:2], c[2:])
        assert np.array_equal(got[k], want), (k, c)


def test_range_counts_empty(dev):
    cov = dev.Coverage(dev.to_device_i32(np.zeros(0, np.int32)), dev.to_device_i32(np.zeros(0, np.int32)))
    assert np.array_equal(cov.range_counts([[1, 2, 3, 4]]), np.zeros((1, 123), np.int32))
    assert cov.range_counts(np.zeros((0, 4))).shape == (0, 123)
