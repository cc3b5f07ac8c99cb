"""GPU parity: libcloops_b200 (through its C ABI) against the golden vectors the reference produced
and against the CPU restatement (oracle/spec.py) on seeded inputs.  Bit-exact: labels, counts and
records are integers."""
import ctypes as C
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


@pytest.fixture(scope="module")
def chr21(gold_dir):
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    return d["X"], d["Y"]


def _labels(dev, X, Y, eps, mp, variant, cut=0):
    dx, dy = dev.to_device_i32(X), dev.to_device_i32(Y)
    lab, info = dev.dbscan_device(dx, dy, eps, mp, variant, cut)
    return lab.cpu().numpy().astype(np.int64), info


def _explain(got, want):
    bad = np.flatnonzero(got != want)
    return "%d/%d rows differ; first rows %s got %s want %s" % (len(bad), len(got), bad[:8], got[bad[:8]], want[bad[:8]])


VAR = {"v1": 1, "v2": 2, "block": 3}
SPEC = {"v1": spec.cdbscan_v1, "v2": spec.cdbscan_v2, "block": spec.blockdbscan}


def test_neighbour_counts_battery(dev, gold_dir):
    bat = np.load(os.path.join(gold_dir, "battery_labels.npz"))
    for c in range(int(bat["ncase"])):
        mat = bat["c%d_mat" % c].astype(np.int64)
        eps, mp = (int(x) for x in bat["c%d_par" % c])
        want = spec.neighbour_counts(mat[:, 1], mat[:, 2], eps)
        dx, dy = dev.to_device_i32(mat[:, 1]), dev.to_device_i32(mat[:, 2])
        got = dev.neighbour_counts_device(dx, dy, eps, 0).cpu().numpy()
        assert np.array_equal(got, want), (c, eps, _explain(got, want))
        got = dev.neighbour_counts_device(dx, dy, eps, mp).cpu().numpy()
        assert np.array_equal(got, np.minimum(want, mp)), (c, eps, mp)


def test_neighbour_counts_chr21(dev, chr21):
    X, Y = chr21
    dx, dy = dev.to_device_i32(X), dev.to_device_i32(Y)
    for eps in (500, 2000, 10000):
        want = spec.neighbour_counts(X, Y, eps)
        got = dev.neighbour_counts_device(dx, dy, eps, 0).cpu().numpy()
        assert np.array_equal(got, want), (eps, _explain(got, want))


@pytest.mark.parametrize("variant", ["v2", "v1", "block"])
def test_battery_labels(dev, gold_dir, variant):
    bat = np.load(os.path.join(gold_dir, "battery_labels.npz"))
    fails = []
    for c in range(int(bat["ncase"])):
        mat = bat["c%d_mat" % c].astype(np.int64)
        eps, mp = (int(x) for x in bat["c%d_par" % c])
        got, _ = _labels(dev, mat[:, 1], mat[:, 2], eps, mp, VAR[variant])
        want = bat["c%d_%s" % (c, variant)].astype(np.int64)
        if not np.array_equal(got, want):
            fails.append((c, eps, mp, _explain(got, want)))
    assert not fails, "%d cases differ: %s" % (len(fails), fails[:5])


@pytest.mark.parametrize("variant", ["v2", "v1", "block"])
@pytest.mark.parametrize("eps,mp", [(500, 5), (1000, 5), (2000, 5), (5000, 20), (2500, 30)])
def test_chr21_labels(dev, gold_dir, chr21, variant, eps, mp):
    X, Y = chr21
    gold = np.load(os.path.join(gold_dir, "chr21_labels.npz"))
    got, info = _labels(dev, X, Y, eps, mp, VAR[variant])
    want = gold["%s_eps%d_mp%d" % (variant, eps, mp)].astype(np.int64)
    assert np.array_equal(got, want), _explain(got, want)
    assert info["n_labelled"] == int((want >= 0).sum())


@pytest.mark.parametrize("variant", ["v2", "v1", "block"])
def test_chr21_cut_rounds(dev, gold_dir, chr21, variant):
    """pipe -m 1 rounds 2 and 3: the cut filter runs on the device (pipe.py:59-63)."""
    X, Y = chr21
    gold = np.load(os.path.join(gold_dir, "chr21_labels.npz"))
    for cut, eps in ((4601, 1000), (13532, 2000)):
        got, info = _labels(dev, X, Y, eps, 5, VAR[variant], cut=cut)
        m = (Y.astype(np.int64) - X) >= cut
        assert info["n_active"] == int(m.sum())
        assert np.all(got[~m] == -1)
        want = gold["%s_cut%d_eps%d_mp5" % (variant, cut, eps)].astype(np.int64)
        assert np.array_equal(got[m], want), _explain(got[m], want)


@pytest.mark.parametrize("variant", ["v2", "v1", "block"])
def test_random_vs_oracle(dev, variant):
    """Seeded inputs larger than the golden battery, checked against oracle/spec.py on the box."""
    from cloops_b200 import synth
    rng = np.random.default_rng(99)
    cases = []
    for n, L, eps, mp in ((20000, 2_000_000, 300, 4), (60000, 5_000_000, 1000, 5), (50000, 400_000, 2000, 20),
                          (30000, 100_000, 50, 3), (40000, 3_000_000, 5000, 10)):
        X, Y = synth.chromosome(n, L, int(rng.integers(1 << 30)), loop_frac=0.2, sigma=eps * 0.7)
        cases.append((X, Y, eps, mp))
    # heavy duplicates and a negative-coordinate case (rotated floor cells straddle u = 0)
    X = rng.integers(0, 500, 20000) * 10
    Y = X + rng.integers(0, 60, 20000) * 10
    cases.append((X.astype(np.int32), Y.astype(np.int32), 25, 6))
    X = rng.integers(-3000, 3000, 8000)
    Y = rng.integers(-3000, 3000, 8000)
    cases.append((X.astype(np.int32), Y.astype(np.int32), 40, 4))
    for k, (X, Y, eps, mp) in enumerate(cases):
        want = SPEC[variant](X.astype(np.int64), Y.astype(np.int64), eps, mp)
        got, _ = _labels(dev, X, Y, eps, mp, VAR[variant])
        assert np.array_equal(got, want), (k, eps, mp, _explain(got, want))


def test_cluster_summary_chr21(dev, gold_dir, chr21):
    X, Y = chr21
    pipe = np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))
    dx, dy = dev.to_device_i32(X), dev.to_device_i32(Y)
    lab, info = dev.dbscan_device(dx, dy, 500, 5, 2)
    bbox, size, kind, row_kind = (t.cpu().numpy() for t in dev.cluster_summary_device(dx, dy, lab, info["n_clusters"]))
    inter = bbox[kind == 1]
    assert np.array_equal(inter, pipe["round0_records"])
    assert int((kind == 2).sum()) == int(pipe["round_nS"][0])
    assert int((row_kind == 1).sum()) == int(pipe["round_ndis"][0])
    assert int((row_kind == 2).sum()) == int(pipe["round_ndss"][0])
    want_i, want_s, in_i, in_s = spec.cluster_records(X, Y, lab.cpu().numpy())
    assert np.array_equal(row_kind == 1, in_i) and np.array_equal(row_kind == 2, in_s)
    assert np.array_equal(size, np.bincount(lab.cpu().numpy()[lab.cpu().numpy() >= 0], minlength=info["n_clusters"]))


def test_host_abi_and_facade(dev, gold_dir, chr21):
    from cloops_b200 import _lib
    from cloops_b200.cDBSCAN2 import cDBSCAN
    X, Y = chr21
    gold = np.load(os.path.join(gold_dir, "chr21_labels.npz"))
    n = len(X)
    mat = np.stack([np.arange(n) * 3 + 7, X, Y], axis=1).astype(np.int64)     # non-contiguous ids
    out = np.empty(n, np.int64)
    info = (C.c_int64 * 8)()
    _lib.check(_lib.lib().cloops_dbscan_host(mat.ctypes.data, n, 1000, 5, 2, out.ctypes.data, C.addressof(info)))
    want = gold["v2_eps1000_mp5"].astype(np.int64)
    assert np.array_equal(out, want)
    db = cDBSCAN(mat, 1000, 5)
    assert db.labels == {int(i): int(c) for i, c in zip(mat[:, 0], want) if c >= 0}
    assert cDBSCAN(np.zeros((0, 3), np.int64), 1000, 5).labels == {}


def test_errors(dev):
    from cloops_b200 import _lib
    dx = dev.to_device_i32(np.array([1, 2, 3], np.int32))
    with pytest.raises(_lib.CloopsError):
        dev.dbscan_device(dx, dx, 0, 5, 2)
    with pytest.raises(_lib.CloopsError):
        dev.dbscan_device(dx, dx, 10, 0, 2)
    with pytest.raises(_lib.CloopsError):
        dev.to_device_i32(np.array([1 << 31], np.int64))
