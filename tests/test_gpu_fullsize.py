"""Bit-exact parity AT FULL SIZE: the CUDA path against the C oracle (oracle/coracle.c, pinned to
the reference's outputs by tests/test_coracle.py) on the inputs BASELINE.json names -- config 2 (10 M ChIA-PET PETs),
config 4's chr1 (16.4 M PETs at Hi-C density, minPts 20-50, eps 5000-10000, incl. a cut-filtered round) and chr21, config 3's
chr1 -- plus the sha256 digests the oracle produced in the build container (tests/golden/fullsize_digests.json)."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import make_digests  # noqa: E402
from oracle.make_digests import sha  # noqa: E402


@pytest.fixture(scope="module")
def digests(gold_dir):
    with open(os.path.join(gold_dir, "fullsize_digests.json")) as fh:
        return json.load(fh)


_INPUTS = {}


def _input(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if name not in _INPUTS:
        _INPUTS.clear()                                   # one full-size input resident at a time
        from cloops_b200 import device
        X, Y = make_digests.CASES[name][0]()
        _INPUTS[name] = (X, Y, device.to_device_i32(X), device.to_device_i32(Y))
    return _INPUTS[name]


CASES = [(name, run, k == 0) for name, (_, runs) in make_digests.CASES.items() for k, run in enumerate(runs)]


@pytest.mark.parametrize("name,run,score", CASES, ids=["%s-v%d_eps%d_mp%d_cut%d" % ((c[0],) + c[1]) for c in CASES])
def test_fullsize_labels_records_counts(digests, name, run, score):
    from cloops_b200 import device
    from oracle import coracle
    variant, eps, minPts, cut = run
    X, Y, dx, dy = _input(name)
    want = digests[name]["v%d_eps%d_mp%d_cut%d" % run]
    assert sha(np.stack([X, Y])) == digests[name]["input"], "synthetic generator drifted"
    lab_t, info = device.dbscan_device(dx, dy, eps, minPts, variant, cut)
    lab = lab_t.cpu().numpy()
    # live C oracle on the same input (seconds), then the committed digest
    X64, Y64 = X.astype(np.int64), Y.astype(np.int64)
    m = (Y64 - X64) >= cut
    ref = np.full(len(X), -1, np.int32)
    ref[m] = coracle.dbscan(X64[m], Y64[m], eps, minPts, variant)
    assert np.array_equal(lab, ref), "labels differ from the C oracle in %d rows" % int((lab != ref).sum())
    assert sha(lab) == want["labels"] and info["n_clusters"] == want["n_clusters"]
    if variant != 3:
        assert info["n_dead"] == want["n_dead"] or variant == 1
    bbox, size, kind, row_kind = device.cluster_summary_device(dx, dy, lab_t, info["n_clusters"])
    bbox_h = bbox.cpu().numpy()
    bbox_h[size.cpu().numpy() == 0] = 0          # ids without members (v1 keeps the gaps of deleted clusters)
    assert sha(bbox_h) == want["bbox"] and sha(kind.cpu().numpy()) == want["kind"]
    if score:
        kind_h = kind.cpu().numpy()
        cand = bbox_h[kind_h == 1].astype(np.int64)
        cand[:, 0] = np.maximum(cand[:, 0], 0)
        cand[:, 2] = np.maximum(cand[:, 2], 0)
        cov = device.Coverage(dx, dy)
        got = cov.range_counts(cand)
        cov.close()
        assert sha(got.astype(np.int32)) == want["counts"]
        pick = np.linspace(0, len(cand) - 1, 300).astype(int)
        assert np.array_equal(got[pick], coracle.range_counts(X64, Y64, cand[pick]))


@pytest.mark.parametrize("name,eps,caps", [("config2_10M", 1000, (5, 0)), ("config4_chr1_16M", 5000, (20, 50)), ("config4_chr1_16M", 10000, (30,))])
def test_fullsize_neighbour_counts(name, eps, caps):
    """The region query itself at full size: min(n(p), cap) for every PET == C oracle."""
    from cloops_b200 import device
    from oracle import coracle
    X, Y, dx, dy = _input(name)
    for cap in caps:
        if cap == 0 and name != "config2_10M":
            continue
        got = device.neighbour_counts_device(dx, dy, eps, cap).cpu().numpy()
        want = coracle.neighbour_counts(X, Y, eps, cap)
        assert np.array_equal(got, want), (name, eps, cap, int((got != want).sum()))
