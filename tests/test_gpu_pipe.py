"""End-to-end drop-in check on the GPU: the ``cLoops -m 1`` pipeline on the reference's bundled chr21
example must reproduce the reference's own ``.loop`` file byte for byte (tests/golden/chr21_m1.loop,
written by the reference through oracle/make_golden.py)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _write_bedpe(path, X, Y):
    # a PET whose two reads are single points: centres cA = X, cB = Y (cLoops/io.py:55-56)
    with open(path, "w") as fh:
        for i, (x, y) in enumerate(zip(X.tolist(), Y.tolist())):
            fh.write("chr21\t%d\t%d\tchr21\t%d\t%d\tp%d\t.\t+\t-\n" % (x, x, y, y, i))


def test_pipe_m1_chr21(need_gpu, gold_dir, tmp_path, monkeypatch):
    from cloops_b200 import pipe
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    gold = np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))
    bedpe = str(tmp_path / "chr21.bedpe")
    _write_bedpe(bedpe, d["X"], d["Y"])
    monkeypatch.chdir(tmp_path)
    cuts = []
    orig = pipe._round

    def spy(fs, eps, minPts, cut, weights=None):
        r = orig(fs, eps, minPts, cut, weights)
        cuts.append((r[2], r[3], r[4]))
        return r

    monkeypatch.setattr(pipe, "_round", spy)
    pipe.pipe([bedpe], "out", [500, 1000, 2000], [5], cpu=1, tmp=1, hic=0, washU=1, juice=1)
    assert [c[2] for c in cuts] == [int(x) for x in gold["round_cut_out"]] == [4601, 13532, 11103]
    assert [c[0] for c in cuts] == [int(x) for x in gold["round_ndis"]]
    assert [c[1] for c in cuts] == [int(x) for x in gold["round_ndss"]]
    got = open(tmp_path / "out.loop", "rb").read()
    want = open(os.path.join(gold_dir, "chr21_m1.loop"), "rb").read()
    assert got == want
    assert os.path.isfile(tmp_path / "out_loops_washU.txt") and os.path.isfile(tmp_path / "out_loops_juicebox.txt")
    assert os.path.isdir(tmp_path / "out")           # -s keeps the .jd directory


def test_run_dbscan_round_records(need_gpu, gold_dir, tmp_path):
    import joblib
    from cloops_b200 import pipe
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    gold = np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))
    n = len(d["X"])
    f = str(tmp_path / "chr21-chr21.jd")
    joblib.dump(np.stack([np.arange(n), d["X"], d["Y"]], axis=1).astype(np.int64), f)
    for k, (eps, cut) in enumerate(((500, 0), (1000, 4601), (2000, 13532))):
        dataI, dataS, dis, dss = pipe.runDBSCAN([f], eps, 5, cut)
        recs = np.array([[r[1], r[2], r[4], r[5]] for r in dataI[("chr21", "chr21")]["records"]])
        assert np.array_equal(recs, gold["round%d_records" % k])
        assert len(dataS) == int(gold["round_nS"][k])
        assert (len(dis), len(dss)) == (int(gold["round_ndis"][k]), int(gold["round_ndss"][k]))
        # host estimate (the reference's numpy path) and the device-side reduction give the same integer cut
        host_cut = pipe.estIntSelCutFrag(np.array(dis), np.array(dss))[0]
        _, _, n_dis, n_dss, dev_cut, _ = pipe._round([f], eps, 5, cut)
        assert (n_dis, n_dss, dev_cut) == (len(dis), len(dss), host_cut) and host_cut == int(gold["round_cut_out"][k])


def test_getintsig_matches_reference_tuples(need_gpu, gold_dir, tmp_path):
    import joblib
    from cloops_b200 import cModel
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    gold = np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))
    n = len(d["X"])
    f = str(tmp_path / "chr21-chr21.jd")
    joblib.dump(np.stack([np.arange(n), d["X"], d["Y"]], axis=1).astype(np.int64), f)
    model, N = cModel.getGenomeCoverage(f)
    assert N == int(gold["sig_N"])
    for k in (0, 7, 100, 555, 800):
        r = gold["sig_records"][k]
        iva, ivb = [max(0, int(r[0])), int(r[1])], [max(0, int(r[2])), int(r[3])]
        got = cModel.getMultiplePsFdr(iva, ivb, model, N)
        assert tuple(float(x) for x in got) == tuple(float(x) for x in gold["sig_tuples"][k][5:])
        assert cModel.getPETsforRegions(iva, ivb, model) == tuple(int(x) for x in gold["sig_tuples"][k][5:8])
    # host-side getCounts keeps the reference's set-of-row-indices contract
    X = d["X"].astype(np.int64)
    assert cModel.getCounts([41000000, 41010000], model[0]) == set(np.flatnonzero((X >= 41000000) & (X <= 41010000)).tolist())


def test_pipe_multi_chrom_hic(need_gpu, gold_dir, tmp_path, monkeypatch):
    """3 chromosomes, 2 eps x 2 minPts rounds, -hic marks: byte-identical to the reference's .loop
    (tests/golden/multi_hic.*, written by oracle/make_golden_multi.py)."""
    from cloops_b200 import pipe
    gold = np.load(os.path.join(gold_dir, "multi_hic.npz"))
    bedpe = str(tmp_path / "in.bedpe")
    with open(bedpe, "w") as fh:
        for name in ("chr1", "chr2", "chrX"):
            for x, y in zip(gold[name + "_X"].tolist(), gold[name + "_Y"].tolist()):
                fh.write("%s\t%d\t%d\t%s\t%d\t%d\tp\t.\t+\t-\n" % (name, x, x, name, y, y))
    monkeypatch.chdir(tmp_path)
    cuts = []
    orig = pipe._round
    monkeypatch.setattr(pipe, "_round", lambda fs, e, m, c, w=None: (lambda r: (cuts.append((r[2], r[3], r[4])), r)[1])(orig(fs, e, m, c, w)))
    pipe.pipe([bedpe], "out", [1000, 2000], [8, 5], cpu=1, tmp=0, hic=1)
    assert np.array_equal(np.array(cuts, np.int64), gold["cuts"])
    assert open(tmp_path / "out.loop", "rb").read() == open(os.path.join(gold_dir, "multi_hic.loop"), "rb").read()
    assert not os.path.isdir(tmp_path / "out")       # temp .jd directory removed without -s


def test_cli_main(need_gpu, gold_dir, tmp_path, monkeypatch):
    """The ``cLoops`` command line (cLoops/pipe.py:298-352): -m 1 and the equivalent explicit lists."""
    from cloops_b200 import pipe
    d = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    bedpe = str(tmp_path / "chr21.bedpe")
    _write_bedpe(bedpe, d["X"], d["Y"])
    monkeypatch.chdir(tmp_path)
    want = open(os.path.join(gold_dir, "chr21_m1.loop"), "rb").read()
    pipe.main(["-f", bedpe, "-o", "a", "-m", "1"])
    assert open(tmp_path / "a.loop", "rb").read() == want
    assert not os.path.isdir(tmp_path / "a")
    pipe.main(["-f", bedpe, "-o", "b", "-eps", "2000,500,1000", "-minPts", "5", "-c", "chr21", "-s"])
    assert open(tmp_path / "b.loop", "rb").read() == want
    assert os.path.isdir(tmp_path / "b")
    pipe.main(["-f", bedpe, "-o", "b", "-m", "1"])           # existing output directory: refuse, leave results alone
    assert open(tmp_path / "b.loop", "rb").read() == want


def test_pipe_auto_eps(need_gpu, gold_dir, tmp_path, monkeypatch):
    """eps = 0: duplicate-dropping ingest + estFragSize (cLoops/pipe.py:229-239, io.py:62-129)."""
    from cloops_b200 import pipe
    gold = np.load(os.path.join(gold_dir, "multi_hic.npz"))
    bedpe = str(tmp_path / "in.bedpe")
    with open(bedpe, "w") as fh:
        for name in ("chr1", "chr2", "chrX"):
            for x, y in zip(gold[name + "_X"].tolist(), gold[name + "_Y"].tolist()):
                fh.write("%s\t%d\t%d\t%s\t%d\t%d\tp\t.\t+\t-\n" % (name, x, x, name, y, y))
    monkeypatch.chdir(tmp_path)
    pipe.pipe([bedpe], "out", 0, [6], cpu=1, tmp=0, hic=0)
    assert open(tmp_path / "out.loop", "rb").read() == open(os.path.join(gold_dir, "multi_auto_eps.loop"), "rb").read()
