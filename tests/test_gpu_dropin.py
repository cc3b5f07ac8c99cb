"""Drop-in test through the REFERENCE'S OWN pipeline: the unmodified cLoops
``pipe()`` (oracle/_ref or /root/reference through oracle/ref_shim.py) runs with ONE name rebound --
``cLoops.pipe.DBSCAN`` (cLoops/pipe.py:42,70) -> the CUDA-backed class -- and must write the same ``.loop`` file, byte
for byte, as the reference did with its own clusterer (tests/golden/chr21_m1.loop)."""
import logging
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_shim  # noqa: E402


@pytest.fixture(scope="module")
def ns():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not ref_shim.available():
        pytest.fail("reference tree not found: oracle/_ref is made by __graft_entry__.build() (oracle/make_ref.py)")
    ns = ref_shim.load()
    ns.pipe.logger = logging.getLogger("ref_dropin")
    ns.pipe.logger.handlers, ns.pipe.logger.propagate = [logging.NullHandler()], False
    return ns


def _write_bedpe(path, X, Y):
    with open(path, "w") as fh:
        for i, (x, y) in enumerate(zip(X.tolist(), Y.tolist())):
            fh.write("chr21\t%d\t%d\tchr21\t%d\t%d\tp%d\t.\t+\t-\n" % (x, x, y, y, i))


def test_reference_pipe_with_cuda_clusterer(ns, gold_dir, tmp_path, monkeypatch):
    from cloops_b200.cDBSCAN2 import cDBSCAN as ours
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    bedpe = str(tmp_path / "chr21.bedpe")
    _write_bedpe(bedpe, d["X"], d["Y"])
    monkeypatch.chdir(tmp_path)
    calls = []

    class Spy(ours):                                   # the class the reference instantiates: DBSCAN(mat, eps, minPts).labels
        def __init__(self, mat, eps, minPts):
            calls.append((len(mat), eps, minPts))
            super().__init__(mat, eps, minPts)

    monkeypatch.setattr(ns.pipe, "DBSCAN", Spy)
    ns.pipe.pipe([bedpe], "out", [500, 1000, 2000], [5], cpu=1, tmp=0, hic=0)
    gold = np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))
    assert [c[1] for c in calls] == [500, 1000, 2000] and calls[0][0] == len(d["X"])
    assert [c[0] for c in calls][1:] == [int((d["Y"].astype(np.int64) - d["X"] >= c).sum()) for c in gold["round_cut_out"][:2]]
    got = open(tmp_path / "out.loop", "rb").read()
    want = open(os.path.join(gold_dir, "chr21_m1.loop"), "rb").read()
    assert got == want


def test_reference_side_scripts_boundary_v1(ns, gold_dir):
    """The v1 class as scripts/jd2saturation:69-70 and scripts/callStripes:51-52 use it: ``pd.Series(DBSCAN(mat, eps,
    minPts).labels)`` on an int64 [N,3] matrix -- labels equal to the reference class's on the same matrix, including
    the anisotropically scaled coordinates of callStripes (:42-43), which leave the int32 range."""
    import pandas as pd
    from cloops_b200.cDBSCAN import cDBSCAN as V1
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    mat = np.stack([np.arange(len(d["X"])), d["X"], d["Y"]], axis=1).astype(np.int64)[:30000]
    want = pd.Series(ns.cDBSCAN(mat, 1000, 5).labels)
    got = pd.Series(V1(mat, 1000, 5).labels)
    assert got.sort_index().equals(want.sort_index())
    stripe = mat.copy()
    stripe[:, 2] = stripe[:, 2] * 50                   # callStripes: horizontal stripes, y scaled by 50 (exceeds 2^30)
    want = pd.Series(ns.cDBSCAN(stripe, 2000, 10).labels)
    got = pd.Series(V1(stripe, 2000, 10).labels)
    assert got.sort_index().equals(want.sort_index())
