"""The CPU restatement (oracle/spec.py) against golden vectors produced by the reference itself
(oracle/make_golden.py, executed in the build container).  CPU only."""
import os

import numpy as np
import pytest

from oracle import spec

VARIANTS = {"v1": spec.cdbscan_v1, "v2": spec.cdbscan_v2, "block": spec.blockdbscan}


@pytest.fixture(scope="module")
def chr21(gold_dir):
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    return d["X"].astype(np.int64), d["Y"].astype(np.int64)


@pytest.fixture(scope="module")
def chr21_labels(gold_dir):
    return np.load(os.path.join(gold_dir, "chr21_labels.npz"))


@pytest.mark.parametrize("variant", ["v2", "v1", "block"])
@pytest.mark.parametrize("eps,mp", [(500, 5), (2000, 5), (5000, 20)])
def test_chr21_labels(chr21, chr21_labels, variant, eps, mp):
    X, Y = chr21
    got = VARIANTS[variant](X, Y, eps, mp)
    want = chr21_labels["%s_eps%d_mp%d" % (variant, eps, mp)]
    assert np.array_equal(got, want)


def test_chr21_counts_appendix_c(chr21_labels):
    # labelled / clusters for eps 500/1000/2000, minPts 5
    want = {"v2": [(15016, 657), (21015, 954), (31966, 1734)],
            "v1": [(15032, 661), (21051, 963), (31998, 1743)],
            "block": [(16158, 753), (22887, 1122), (36772, 2142)]}
    for v, rows in want.items():
        for eps, (nl, nc) in zip((500, 1000, 2000), rows):
            lab = chr21_labels["%s_eps%d_mp5" % (v, eps)]
            assert int((lab >= 0).sum()) == nl
            assert len(np.unique(lab[lab >= 0])) == nc


@pytest.mark.parametrize("variant", ["v2", "v1", "block"])
def test_chr21_cut_filtered(chr21, chr21_labels, variant):
    X, Y = chr21
    for cut, eps in ((4601, 1000), (13532, 2000)):
        m = (Y - X) >= cut
        got = VARIANTS[variant](X[m], Y[m], eps, 5)
        assert np.array_equal(got, chr21_labels["%s_cut%d_eps%d_mp5" % (variant, cut, eps)])


@pytest.mark.parametrize("variant", ["v2", "v1", "block"])
def test_battery(gold_dir, variant):
    bat = np.load(os.path.join(gold_dir, "battery_labels.npz"))
    for c in range(int(bat["ncase"])):
        mat = bat["c%d_mat" % c].astype(np.int64)
        eps, mp = (int(x) for x in bat["c%d_par" % c])
        got = VARIANTS[variant](mat[:, 1], mat[:, 2], eps, mp)
        assert np.array_equal(got, bat["c%d_%s" % (c, variant)]), (c, eps, mp)


def test_records_round1(chr21, chr21_labels, gold_dir):
    X, Y = chr21
    pipe = np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))
    inter, selfl, in_i, in_s = spec.cluster_records(X, Y, chr21_labels["v2_eps500_mp5"])
    assert np.array_equal(inter[:, :4], pipe["round0_records"])
    assert len(selfl) == int(pipe["round_nS"][0])
    assert int(in_i.sum()) == int(pipe["round_ndis"][0]) == 4867
    assert int(in_s.sum()) == int(pipe["round_ndss"][0]) == 10149


def test_range_counts_and_stats(chr21, gold_dir):
    X, Y = chr21
    pipe = np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))
    recs, ints, tup = pipe["sig_records"], pipe["sig_ints200"], pipe["sig_tuples"]
    N = int(pipe["sig_N"])
    assert N == len(X)
    for k in range(0, 200, 5):
        r = recs[k]
        iva = [max(0, int(r[0])), int(r[1])]
        ivb = [max(0, int(r[2])), int(r[3])]
        c = spec.range_counts(X, Y, iva, ivb)
        assert np.array_equal(c, ints[k]), k
        got = spec.stats_from_counts(c, N)
        want = tup[k][5:]
        assert tuple(float(x) for x in got) == tuple(float(x) for x in want), k
