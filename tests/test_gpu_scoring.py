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


@pytest.mark.parametrize("variant", [2, 1, 3])
def test_fused_pass_equals_separate_calls(dev, gold_dir, variant):
    """cloops_pass_run (one C-ABI call, coverage on a side stream) against the separate entry points."""
    from cloops_b200 import hotpath
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    dx, dy = dev.to_device_i32(d["X"]), dev.to_device_i32(d["Y"])
    for eps, mp, cut in ((500, 5, 0), (1000, 5, 4601), (5000, 20, 0)):
        ref = hotpath.run_device_separate(dx, dy, eps, mp, variant, cut, rows=True)
        p = dev.Pass(dx, dy, eps, mp, variant, cut)
        assert p.info == ref.info and p.n_clusters == ref.info["n_clusters"]
        assert torch.equal(p.bbox, ref.bbox) and torch.equal(p.size, ref.size) and torch.equal(p.kind, ref.kind)
        assert torch.equal(p.cand, ref.cand) and torch.equal(p.counts, ref.counts)
        assert int((p.member_kind == 1).sum()) == int((ref.row_kind == 1).sum())
        assert int((p.member_kind == 2).sum()) == int((ref.row_kind == 2).sum())
        # members are a permutation of the clustered rows
        lab = ref.labels
        if variant != 3:
            act = (dy - dx) >= cut if cut > 0 else torch.ones_like(dx, dtype=torch.bool)
            key = lambda x, y, l: (x.long() * 1_000_003 + y.long() * 7 + l.long()).sort().values
            assert torch.equal(key(p.xs, p.ys, p.labels_sorted), key(dx[act], dy[act], lab[act]))
        p.close()
    # host entry + fetch
    hx, hy = torch.from_numpy(d["X"]).pin_memory(), torch.from_numpy(d["Y"]).pin_memory()
    p = dev.Pass(hx, hy, 500, 5, variant, 0, host=True)
    hb = torch.empty((p.n_clusters, 4), dtype=torch.int32).pin_memory()
    hk = torch.empty(p.n_clusters, dtype=torch.uint8).pin_memory()
    hm = torch.empty(p.n_members, dtype=torch.uint8).pin_memory()
    hc = torch.empty((p.n_candidates, 123), dtype=torch.int32).pin_memory()
    p.fetch(hb, hk, hm, hc)
    ref = hotpath.run_device_separate(dx, dy, 500, 5, variant, 0)
    assert torch.equal(hb, ref.bbox.cpu()) and torch.equal(hk, ref.kind.cpu()) and torch.equal(hc, ref.counts.cpu())
    p.close()
    # empty input and no-candidate input
    e = dev.Pass(dx[:0], dy[:0], 500, 5, variant)
    assert e.n_clusters == 0 and e.n_candidates == 0 and e.counts.shape == (0, 123)
    e.close()
    q = dev.Pass(dx[:50], dy[:50], 5, 40, variant)
    assert q.n_clusters == 0 and q.n_candidates == 0
    q.close()
