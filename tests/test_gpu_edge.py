"""Edge cases through the C ABI against the CPU oracle: tiny inputs, degenerate geometry, extreme
parameters, refusals.  Bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import spec  # noqa: E402

SPEC = {1: spec.cdbscan_v1, 2: spec.cdbscan_v2, 3: spec.blockdbscan}


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cloops_b200 import device
    return device


def _check(dev, X, Y, eps, mp, variants=(1, 2, 3), cut=0):
    X = np.asarray(X, np.int64)
    Y = np.asarray(Y, np.int64)
    for v in variants:
        got, info = dev.dbscan_device(dev.to_device_i32(X), dev.to_device_i32(Y), eps, mp, v, cut)
        got = got.cpu().numpy()
        if cut > 0:
            m = (Y - X) >= cut
            want = np.full(len(X), -1, np.int64)
            want[m] = SPEC[v](X[m], Y[m], eps, mp)
        else:
            want = SPEC[v](X, Y, eps, mp)
        assert np.array_equal(got, want), (v, eps, mp, got[:10], want[:10])


def test_tiny_inputs(dev):
    _check(dev, [5], [9], 10, 1)
    _check(dev, [5], [9], 10, 2)
    _check(dev, [5, 6], [9, 9], 10, 2)
    _check(dev, [5, 600], [9, 900], 10, 1)
    _check(dev, [0, 0, 0], [0, 0, 0], 1, 3)


def test_all_identical_points(dev):
    n = 3000
    _check(dev, np.full(n, 1234), np.full(n, 5678), 7, 5)
    _check(dev, np.full(n, 1234), np.full(n, 5678), 7, n)          # exactly minPts points
    _check(dev, np.full(n, 1234), np.full(n, 5678), 7, n + 1)      # one short: no cluster at all


def test_minpts_one_and_eps_one(dev):
    rng = np.random.default_rng(1)
    X = rng.integers(0, 400, 5000)
    Y = X + rng.integers(0, 400, 5000)
    _check(dev, X, Y, 1, 1)
    _check(dev, X, Y, 1, 2)
    _check(dev, X, Y, 3, 1)
    _check(dev, X, Y, 2, 4)


def test_huge_eps_single_strip(dev):
    rng = np.random.default_rng(2)
    X = rng.integers(0, 5000, 4000)
    Y = X + rng.integers(0, 5000, 4000)
    _check(dev, X, Y, 1 << 20, 5)          # everything within eps of everything
    _check(dev, X, Y, 20000, 4001)         # minPts above n
    _check(dev, X, Y, 3000, 50)            # long strips: exercises the global-memory fallback of the region query


def test_long_dense_strips_fallback(dev):
    """Strips far longer than the shared-memory tile (dense diagonal): the region query must agree
    between its tiled and its fallback path."""
    rng = np.random.default_rng(3)
    n = 60000
    X = rng.integers(0, 3_000_000, n)
    Y = X + rng.integers(0, 40, n)          # all PETs hug the diagonal: u in [-40, 0], strips hold thousands
    for eps, mp in ((5000, 20), (500, 5), (100000, 200)):
        dx, dy = dev.to_device_i32(X), dev.to_device_i32(Y)
        got = dev.neighbour_counts_device(dx, dy, eps, mp).cpu().numpy()
        want = np.minimum(spec.neighbour_counts(X, Y, eps), mp) if eps <= 5000 else None
        if want is not None:
            assert np.array_equal(got, want), (eps, mp)
    _check(dev, X[:20000], Y[:20000], 500, 5)


def test_cut_edge_cases(dev):
    rng = np.random.default_rng(4)
    X = rng.integers(0, 100000, 6000)
    Y = X + rng.integers(0, 3000, 6000)
    _check(dev, X, Y, 300, 4, cut=1500)
    _check(dev, X, Y, 300, 4, cut=1)
    _check(dev, X, Y, 300, 4, cut=10**6)        # removes every row
    _check(dev, X, Y, 300, 4, cut=2999)         # leaves a handful


def test_negative_and_mixed_coordinates(dev):
    rng = np.random.default_rng(5)
    X = rng.integers(-50000, 50000, 8000)
    Y = rng.integers(-50000, 50000, 8000)       # Y < X allowed for generic callers (callStripes-style input)
    _check(dev, X, Y, 700, 4)
    _check(dev, X - 10**8, Y + 10**8, 700, 4)


def test_refusals(dev):
    from cloops_b200._lib import CloopsError
    x = torch.tensor([0, (1 << 30) - 1], dtype=torch.int32, device="cuda")
    with pytest.raises(CloopsError, match="strips"):
        dev.dbscan_device(x, x.clone(), 1, 2, 2)                 # eps 1 over a 2^31 span: > 2^27 strips
    bad = torch.tensor([0, 1 << 30], dtype=torch.int32, device="cuda")
    for v in (1, 2, 3):
        with pytest.raises(CloopsError):
            dev.dbscan_device(bad, bad.clone(), 100, 2, v)
    with pytest.raises(CloopsError):
        dev.dbscan_device(x, x.clone(), 100, 2, 7)               # unknown variant
    # the context stays healthy after refusals
    _check(dev, [1, 2, 3], [4, 5, 6], 10, 2)


def test_index_order_counting_sort_equals_radix_sort(dev, monkeypatch):
    """index_build has two ways to reach the (strip, u', row) order: the counting sort by strip (histogram + arrival
    ranks, windowed scatter, in-strip rank) and the stable radix sort it falls back to for long strips.  Both must
    produce the same index: coordinates in index order, neighbour counts and labels, with and without the cut
    filter, on ties-heavy input too."""
    from cloops_b200 import synth
    rng = np.random.default_rng(5)
    sets = []
    X, Y = synth.chromosome(300_000, 8_000_000, seed=3, loop_frac=0.1, sigma=400.0)
    sets.append((X, Y, 1000, 5, 0))
    sets.append((X, Y, 1000, 5, 3000))
    Xd = (rng.integers(0, 3000, 200_000) * 64).astype(np.int32)            # heavy duplicates / ties in (strip, u')
    Yd = Xd + (rng.integers(0, 50, 200_000) * 64).astype(np.int32)
    sets.append((Xd, Yd, 500, 4, 0))
    for X, Y, eps, mp, cut in sets:
        dx, dy = dev.to_device_i32(X), dev.to_device_i32(Y)
        out = []
        for mode in (None, "radix"):
            if mode:
                monkeypatch.setenv("CLOOPS_INDEX_SORT", mode)
            else:
                monkeypatch.delenv("CLOOPS_INDEX_SORT", raising=False)
            ix = dev.Index(dx, dy, eps, cut)
            xs, ys = ix.coords()
            cnt = ix.count(mp)
            lab, ls, info = ix.dbscan(mp, 2, want_sorted=True)
            n = ix.n_active
            out.append((xs[:n].cpu().numpy(), ys[:n].cpu().numpy(), cnt[:n].cpu().numpy(), lab.cpu().numpy(), ls.cpu().numpy()))
            ix.close()
        monkeypatch.delenv("CLOOPS_INDEX_SORT", raising=False)
        for a, b in zip(*out):
            assert np.array_equal(a, b)


def test_region_query_tile_boundaries_and_unaligned_output(dev):
    """The tiled region query owns 1024 sorted points per CTA (4 per thread, 128-bit stores): sizes around the tile and
    vector boundaries, every templated cap, and an output pointer that is not 16-byte aligned."""
    rng = np.random.default_rng(9)
    for n in (1, 3, 4, 5, 1023, 1024, 1025, 2047, 2048, 2049, 4100):
        X = rng.integers(0, 40 * max(n, 8), n)
        Y = X + rng.integers(0, 3000, n)
        eps = 700
        want = spec.neighbour_counts(X.astype(np.int64), Y.astype(np.int64), eps)
        dx, dy = dev.to_device_i32(X), dev.to_device_i32(Y)
        ix = dev.Index(dx, dy, eps)
        row_of = None
        for cap in (0, 1, 2, 3, 5, 6, 9, 10, 17):
            out = ix.count(cap)[:ix.n_active].cpu().numpy()
            buf = torch.full((ix.n_active + 5,), -3, dtype=torch.int32, device=dx.device)
            ix.count(cap, buf[1:])                                   # 4-byte aligned only
            assert np.array_equal(buf[1:1 + ix.n_active].cpu().numpy(), out)
            assert int(buf[0]) == -3 and int(buf[1 + ix.n_active]) == -3
            # index order -> row order through the coordinates (duplicates share their count, so any matching row will do)
            if row_of is None:
                xs, ys = ix.coords()
                key = {}
                for r, (a, b) in enumerate(zip(X.tolist(), Y.tolist())):
                    key.setdefault((a, b), r)
                row_of = np.array([key[(a, b)] for a, b in zip(xs.cpu().numpy().tolist(), ys.cpu().numpy().tolist())])
            exp = want[row_of] if cap == 0 else np.minimum(want[row_of], cap)
            assert np.array_equal(out, exp), (n, cap)
        ix.close()
