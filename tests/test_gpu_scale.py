"""Full-size (BASELINE.json configs[1]: 10 M PETs) checks through size-independent properties; the
oracle cannot run at this size in seconds, so these complement the bit-exact small-size tests."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N = 10_000_000
EPS, MP = 1000, 5


@pytest.fixture(scope="module")
def big():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cloops_b200 import device, synth
    X, Y = synth.config2(N)
    return device, X, Y, device.to_device_i32(X), device.to_device_i32(Y)


@pytest.mark.parametrize("variant", [2, 1, 3])
def test_cut_filter_equals_prefiltered_input(big, variant):
    """pipe.py:59-63: clustering with cut on the device == clustering the filtered matrix (ids and all)."""
    device, X, Y, dx, dy = big
    cut = 4601
    lab_full, info = device.dbscan_device(dx, dy, EPS, MP, variant, cut)
    m = (Y.astype(np.int64) - X) >= cut
    lab_sub, info_sub = device.dbscan_device(device.to_device_i32(X[m]), device.to_device_i32(Y[m]), EPS, MP, variant)
    lab_full = lab_full.cpu().numpy()
    assert info["n_active"] == int(m.sum()) == info_sub["n_active"]
    assert np.all(lab_full[~m] == -1)
    assert np.array_equal(lab_full[m], lab_sub.cpu().numpy())
    assert info["n_clusters"] == info_sub["n_clusters"]


@pytest.mark.parametrize("variant", [2, 1])
def test_translation_by_cell_multiples(big, variant):
    """Shifting X and Y by eps moves v by 2 eps and leaves u alone: the reference's floor cells
    (cDBSCAN2.py:69-70) shift by whole cells, so labels must be identical."""
    device, X, Y, dx, dy = big
    a, _ = device.dbscan_device(dx, dy, EPS, MP, variant)
    b, _ = device.dbscan_device(dx + 7 * EPS, dy + 7 * EPS, EPS, MP, variant)
    assert torch.equal(a, b)


def test_core_points_and_counts_consistent(big):
    device, X, Y, dx, dy = big
    cnt = device.neighbour_counts_device(dx, dy, EPS, MP)
    full = device.neighbour_counts_device(dx, dy, EPS, 0)
    assert torch.equal(cnt, torch.clamp(full, max=MP))
    core = cnt >= MP
    lab1, info1 = device.dbscan_device(dx, dy, EPS, MP, 1)
    lab2, info2 = device.dbscan_device(dx, dy, EPS, MP, 2)
    assert info1["n_core"] == info2["n_core"] == int(core.sum())
    assert info1["n_components"] == info2["n_components"]
    # v1: every core point is labelled unless its whole cluster was deleted for size (never a core-heavy one)
    assert int((lab1[core] < 0).sum()) <= MP * info1["n_components"]
    # v2: unlabelled core points belong to released clusters, each with < minPts core points
    assert int((lab2[core] < 0).sum()) <= (MP - 1) * info2["n_dead"]
    # v2 ids are dense
    k = info2["n_clusters"]
    assert int(lab2.max()) == k - 1 and len(torch.unique(lab2[lab2 >= 0])) == k
    # an isolated point (count 1) is never labelled
    assert bool((lab2[full == 1] == -1).all()) and bool((lab1[full == 1] == -1).all())


def test_sample_window_matches_oracle(big):
    """A same-density window of the big set, clustered alone, against the CPU oracle."""
    from oracle import spec
    device, X, Y, dx, dy = big
    m = X < 2_000_000
    xs, ys = X[m], Y[m]
    for variant, fn in ((2, spec.cdbscan_v2), (1, spec.cdbscan_v1), (3, spec.blockdbscan)):
        got, _ = device.dbscan_device(device.to_device_i32(xs), device.to_device_i32(ys), EPS, MP, variant)
        want = fn(xs.astype(np.int64), ys.astype(np.int64), EPS, MP)
        assert np.array_equal(got.cpu().numpy(), want), variant


def test_index_reuse_sweep_matches_fresh_runs():
    """BASELINE.json configs[4] (eps x minPts sweep): one resident index per eps serves every minPts
    and both strip-index variants; results must equal independent runs."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cloops_b200 import device, synth
    X, Y = synth.config2(2_000_000, seed=77)
    dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
    for eps in (500, 2500, 10000):
        ix = device.Index(dx, dy, eps)
        try:
            for mp in (50, 5, 20):
                for variant in (2, 1):
                    lab, ls, info = ix.dbscan(mp, variant, want_sorted=True)
                    ref, rinfo = device.dbscan_device(dx, dy, eps, mp, variant)
                    assert torch.equal(lab, ref), (eps, mp, variant)
                    assert info == rinfo
                    xs, ys = ix.coords()
                    # index order is a permutation of the rows: same multiset of (x, y, label)
                    a = torch.stack([xs.long(), ys.long(), ls.long()], 1)
                    b = torch.stack([dx.long(), dy.long(), lab.long()], 1)
                    key = lambda t: (t[:, 0] * 1_000_003 + t[:, 1] * 7 + t[:, 2]).sort().values
                    assert torch.equal(key(a), key(b))
        finally:
            ix.close()
