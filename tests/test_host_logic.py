"""CPU tests of the host side: statistics tail, loop de-duplication, file formats, cut estimation,
chromosome sharding (world_size 2 over gloo).  No GPU, no reference tree needed (golden vectors)."""
import gzip
import logging
import os
import subprocess
import sys

import joblib
import numpy as np
import pandas as pd
import pytest

from cloops_b200 import cModel, dist, ests, io, pipe
from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold(gold_dir):
    return np.load(os.path.join(gold_dir, "chr21_m1_pipe.npz"))


def test_stats_tail_matches_reference_tuples(gold):
    N = int(gold["sig_N"])
    for k in range(200):
        got = cModel._stats(gold["sig_ints200"][k].astype(np.int32), N)
        want = gold["sig_tuples"][k][5:]
        assert tuple(float(x) for x in got) == tuple(float(x) for x in want), k


def test_stats_batch_equals_scalar(gold):
    N = int(gold["sig_N"])
    c = gold["sig_ints200"].astype(np.int32)
    batch = cModel._stats_batch(c, N)
    for k in range(200):
        one = cModel._stats(c[k], N)
        assert tuple(float(b[k]) for b in batch) == tuple(float(x) for x in one), k
        assert tuple(float(b[k]) for b in batch) == tuple(float(x) for x in gold["sig_tuples"][k][5:]), k


def test_scoring_tail_reproduces_loop_file(gold, gold_dir, tmp_path):
    """From the reference's per-candidate tuples, the host tail (key numbering, removeDup x2,
    Bonferroni, markIntSig, to_csv) must reproduce the reference's .loop byte for byte."""
    tup = gold["sig_tuples"]
    ds, i = {}, 0
    for t in tup:
        a0, a1, b0, b1 = (int(x) for x in t[:4])
        ra, rb, rab = int(t[5]), int(t[6]), int(t[7])
        if rab < 5:
            continue
        key = "chr21-chr21-%d" % i
        i += 1
        ds[key] = {"distance": abs((b0 + b1) / 2.0 - (a0 + a1) / 2.0), "ra": ra, "rb": rb, "rab": rab, "ES": t[8], "FDR": t[9],
                   "hypergeometric_p-value": t[10], "poisson_p-value": t[11], "binomial_p-value": t[12],
                   "iva": "chr21:%d-%d" % (a0, a1), "ivb": "chr21:%d-%d" % (b0, b1)}
    ds = cModel.removeDup(cModel.removeDup(ds))
    tab = pd.DataFrame(ds).T
    tab["poisson_p-value_corrected"] = cModel.getBonPvalues(tab["poisson_p-value"])
    tab["binomial_p-value_corrected"] = cModel.getBonPvalues(tab["binomial_p-value"])
    tab["hypergeometric_p-value_corrected"] = cModel.getBonPvalues(tab["hypergeometric_p-value"])
    tab = cModel.markIntSig(tab)
    out = tmp_path / "t.loop"
    tab.to_csv(out, sep="\t", index_label="loopId")
    assert open(out, "rb").read() == open(os.path.join(gold_dir, "chr21_m1.loop"), "rb").read()
    assert int(tab["significant"].sum()) == 202 and len(tab) == 343


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_removedup_and_marks_vs_live_reference():
    ns = ref_shim.load()
    rng = np.random.default_rng(11)
    for trial in range(20):
        n = int(rng.integers(1, 120))
        ds = {}
        for k in range(n):
            a0 = int(rng.integers(0, 3000)); a1 = a0 + int(rng.integers(0, 300))
            b0 = a0 + int(rng.integers(200, 3000)); b1 = b0 + int(rng.integers(0, 300))
            c = "chr%d" % rng.integers(1, 3)
            ds["%s-%s-%d" % (c, c, k)] = {"iva": "%s:%d-%d" % (c, a0, a1), "ivb": "%s:%d-%d" % (c, b0, b1),
                                          "rab": int(rng.integers(1, 30)), "ra": int(rng.integers(30, 90)), "rb": int(rng.integers(30, 90)),
                                          "binomial_p-value": float(10 ** -rng.uniform(2, 9)), "ES": float(rng.uniform(0, 5)),
                                          "FDR": float(rng.choice([0.0, 0.01, 0.02])), "poisson_p-value": float(10 ** -rng.uniform(2, 9)),
                                          "hypergeometric_p-value": float(10 ** -rng.uniform(5, 15))}
        got, want = cModel.removeDup(dict(ds)), ns.cModel.removeDup(dict(ds))
        assert list(got.keys()) == list(want.keys()), trial
        if len(want):
            a, b = pd.DataFrame(got).T, pd.DataFrame(want).T
            assert cModel.markIntSig(a.copy())["significant"].tolist() == ns.cModel.markIntSig(b.copy())["significant"].tolist()
            assert cModel.markIntSigHic(a.copy())["significant"].tolist() == ns.cModel.markIntSigHic(b.copy())["significant"].tolist()


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_removedup_stress_ties_chains_vs_live_reference():
    """cloops_remove_dup against the reference's removeDup (cModel.py:198-259, run live) where its order dependence shows:
    anchors on a coarse lattice (long chains of overlapping loops, a member can overlap a group's leader but not its other
    members), densities and p-values from small sets (shared maxima inside a group, p exactly at the cut-off), reversed
    intervals (the sweep's precondition fails: pair-loop path), and the second pass over the survivors (cModel.py:318,322)."""
    ns = ref_shim.load()
    rng = np.random.default_rng(2024)
    for trial in range(40):
        n = int(rng.integers(2, 260))
        step = int(rng.choice([50, 200, 1000]))
        width = int(rng.choice([1, 2, 4])) * step
        ds = {}
        for k in range(n):
            a0 = int(rng.integers(0, 40)) * step
            a1 = a0 + int(rng.integers(0, 3)) * width // 2
            b0 = a0 + int(rng.integers(2, 30)) * step
            b1 = b0 + int(rng.integers(0, 3)) * width // 2
            if trial % 8 == 7 and rng.random() < 0.1:
                a0, a1 = a1, a0                                     # reversed interval
            ra, rb = int(rng.choice([20, 40, 80])), int(rng.choice([20, 40, 80]))
            ds["chr1-chr1-%d" % k] = {"iva": "chr1:%d-%d" % (a0, a1), "ivb": "chr1:%d-%d" % (b0, b1),
                                      "rab": int(rng.choice([4, 8, 16])), "ra": ra, "rb": rb,
                                      "binomial_p-value": float(rng.choice([1e-6, 1e-5, 2e-5, 1e-9])), "ES": 3.0, "FDR": 0.0,
                                      "poisson_p-value": 1e-7, "hypergeometric_p-value": 1e-12}
        got, want = cModel.removeDup(dict(ds)), ns.cModel.removeDup(dict(ds))
        assert list(got.keys()) == list(want.keys()), trial
        got2, want2 = cModel.removeDup(dict(got)), ns.cModel.removeDup(dict(want))
        assert list(got2.keys()) == list(want2.keys()), trial


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_combine_rounds_vs_live_reference():
    """cloops_combine_rounds (all rounds of a chromosome in one hash pass) against the reference's combineTwice applied round
    after round (cLoops/pipe.py:155-174, run live): boxes repeated inside one round stay (the known set is taken before the
    round is appended), boxes seen in ANY earlier round go; order = round order, then position."""
    ns = ref_shim.load()
    rng = np.random.default_rng(77)
    for trial in range(40):
        n_rounds = int(rng.integers(1, 9))
        pool = rng.integers(0, 50, (int(rng.integers(1, 40)), 4)).astype(np.int32) * np.int32(100)      # few distinct boxes: many repeats
        rounds = [pool[rng.integers(0, len(pool), int(rng.integers(0, 30)))] for _ in range(n_rounds)]
        rounds = [r for r in rounds if len(r)] or [pool[:1]]
        got = pipe._combine_rounds([np.ascontiguousarray(r) for r in rounds])
        dataI = {}
        for r in rounds:
            recs = [["c", int(b[0]), int(b[1]), "c", int(b[2]), int(b[3])] for b in r]
            dataI = ns.pipe.combineTwice(dataI, {("c", "c"): {"f": "f", "records": recs}})
        want = [[r[1], r[2], r[4], r[5]] for r in dataI[("c", "c")]["records"]]
        assert got.tolist() == want, trial


def test_table_from_counts_equals_reference_tail(gold):
    """The columnar statistics tail (tableFromCounts) against the reference's own tail (cModel.py:295-331: dict of dicts,
    removeDup twice, DataFrame(ds).T, Bonferroni) on the same counted integers: identical CSV text."""
    ns = ref_shim.load()
    rng = np.random.default_rng(3)
    N = int(gold["sig_N"])
    for trial in range(6):
        K = int(rng.integers(1, 400))
        counts = gold["sig_ints200"][rng.integers(0, 200, K)].astype(np.int32)
        a0 = rng.integers(0, 40000, K); a1 = a0 + rng.integers(0, 1500, K)
        b0 = a0 + rng.integers(3000, 30000, K); b1 = b0 + rng.integers(0, 1500, K)
        cand = np.stack([a0, a1, b0, b1], axis=1).astype(np.int64)
        keep = np.sort(rng.choice(K, max(1, K // 2), replace=False))
        dist = np.abs((cand[:, 2] + cand[:, 3]) / 2.0 - (cand[:, 0] + cand[:, 1]) / 2.0)
        got = cModel.tableFromCounts({"N": N, "names": ("chr7", "chr7"), "cand": cand, "keep": keep, "dist": dist, "counts": counts[keep]})
        ds = {}
        for i, k in enumerate(keep.tolist()):
            ra, rb, rab, es, fdr, hyp, pop, nbp = cModel._stats(counts[k], N)
            ds["chr7-chr7-%d" % i] = {"distance": float(dist[k]), "ra": ra, "rb": rb, "rab": rab, "ES": es, "FDR": fdr, "hypergeometric_p-value": hyp,
                                      "poisson_p-value": pop, "binomial_p-value": nbp, "iva": "chr7:%d-%d" % (cand[k, 0], cand[k, 1]),
                                      "ivb": "chr7:%d-%d" % (cand[k, 2], cand[k, 3])}
        ds = ns.cModel.removeDup(ds)
        ds = ns.cModel.removeDup(ds) if len(ds) else ds
        if len(ds) == 0:
            assert got is None
            continue
        want = pd.DataFrame(ds).T
        for col in ("poisson_p-value", "binomial_p-value", "hypergeometric_p-value"):
            want[col + "_corrected"] = ns.cModel.getBonPvalues(want[col])
        assert got.to_csv(sep="\t", index_label="loopId") == want.to_csv(sep="\t", index_label="loopId"), trial
        assert ns.cModel.markIntSigHic(got.copy())["significant"].tolist() == ns.cModel.markIntSigHic(want.copy())["significant"].tolist()


def test_nearby_windows_and_overlap():
    ivas, ivbs = cModel.getNearbyPairRegions([100, 301], [1000, 1400])
    assert len(ivas) == len(ivbs) == 10
    assert ivas[0] == [0, 0] and ivas[4] == [0, 150] and ivas[5] == [250, 450]        # clamped at 0 (cModel.py:98-102)
    assert cModel.checkOverlap(["c", 1, 5], ["c", 10, 20], ["c", 5, 9], ["c", 20, 30]) is True
    assert cModel.checkOverlap(["c", 1, 5], ["c", 10, 20], ["d", 5, 9], ["c", 20, 30]) is None
    assert cModel.checkOverlap(["c", 1, 5], ["c", 10, 20], ["c", 6, 9], ["c", 20, 30]) is False


def test_bedpe_ingest_and_jd(tmp_path):
    lines = ["chr1\t100\t200\tchr1\t1000\t1101\tp0\t.\t+\t-",       # cA=150, cB=1050
             "chr1\t5000\t5003\tchr1\t10\t20\tp1\t.\t+\t+",         # swapped: cA=15, cB=5001
             "chr1\t1\t2\tchr2\t3\t4\tp2\t.\t+\t-",                 # trans: dropped
             "chr2\t7\t8\tchr2\t9\t12\tp3\t.\t-\t-",
             "*\t-1\t-1\t*\t-1\t-1\tp4\t.\t+\t-", "short\tline",
             "chr1\t100\t200\tchr1\t1000\t1101\tdup\t.\t+\t-"]
    f = tmp_path / "a.bedpe.gz"
    with gzip.open(f, "wt") as fh:
        fh.write("\n".join(lines) + "\n")
    log = logging.getLogger("t")
    out = tmp_path / "o"
    os.mkdir(out)
    cfs = io.parseRawBedpe2([str(f)], str(out), [], 0, log)
    assert [os.path.basename(c) for c in cfs] == ["chr1-chr1.jd", "chr2-chr2.jd"]
    key, mat = io.parseJd(cfs[0])
    assert key == ("chr1", "chr1") and mat.dtype == np.int64
    assert mat.tolist() == [[0, 150, 1050], [1, 15, 5001], [2, 150, 1050]]
    assert io.parseJd(cfs[0], cut=1000)[1].tolist() == [[1, 15, 5001]]
    out2 = tmp_path / "o2"
    os.mkdir(out2)
    cfs2, dsts = io.parseRawBedpe([str(f)], str(out2), {"chr1"}, 0, log)
    assert io.parseJd(cfs2[0])[1].tolist() == [[0, 150, 1050], [1, 15, 5001]] and dsts == [900]
    txt = tmp_path / "chr9-chr9.txt"
    txt.write_text("0\t5\t9\n1\t6\t10\n")
    assert joblib.load(io.txt2jd(str(txt))).tolist() == [[0, 5, 9], [1, 6, 10]] and not txt.exists()
    assert io.parseIv("chr21:44800894-44801696") == ["chr21", 44800894, 44801696]


def test_cut_estimation_and_round_helpers(gold):
    rng = np.random.default_rng(0)
    di = rng.integers(5000, 500000, 4000).astype(float)
    dsv = rng.integers(50, 3000, 6000).astype(float)
    cut, frag = ests.estIntSelCutFrag(di, dsv)
    lds, ldi = np.log2(dsv), np.log2(di)
    want = int(2 ** min(np.median(lds) + 3 * lds.std(), (lds.mean() * lds.std() + ldi.mean() * ldi.std()) / (lds.std() + ldi.std())))
    assert cut == want and frag == int(2 ** np.median(lds))
    assert ests.estFragSize([100] * 5 + [200] * 3 + [300]) == 200
    if ref_shim.available():
        ns = ref_shim.load()
        assert ns.ests.estIntSelCutFrag(di, dsv) == (cut, frag)
    a = {("c", "c"): {"f": "f", "records": [["c", 1, 2, "c", 30, 40], ["c", 5, 6, "c", 7, 9]]}}
    b = {("c", "c"): {"f": "f", "records": [["c", 1, 2, "c", 30, 40], ["c", 1, 3, "c", 30, 40]]},
         ("d", "d"): {"f": "g", "records": [["d", 1, 2, "d", 3, 4]]}}
    m = pipe.combineTwice(a, b)
    assert [r[1:3] + r[4:6] for r in m[("c", "c")]["records"]] == [[1, 2, 30, 40], [5, 6, 7, 9], [1, 3, 30, 40]]
    assert ("d", "d") in m
    f = pipe.filterClusterByDis(m, 20)
    assert [r[1:3] for r in f[("c", "c")]["records"]] == [[1, 2], [1, 3]]
    assert pipe._int_list("2000,500,1000", False) == [500, 1000, 2000] and pipe._int_list("20,50", True) == [50, 20]
    assert pipe._int_list(0, False) == 0 and pipe._int_list("7", True) == [7]


def test_lpt_assignment():
    owner = dist.assign(list("abcdefgh"), [13, 12, 10, 9, 8, 7, 5, 1], nranks=4)
    loads = [sum(w for w, o in zip([13, 12, 10, 9, 8, 7, 5, 1], owner) if o == r) for r in range(4)]
    assert max(loads) - min(loads) <= 3 and sorted(set(owner)) == [0, 1, 2, 3]
    assert dist.my_share(["x", "y"]) == ["x", "y"] and dist.merge_in_order(["x"], {"x": 1}) == [1]


def test_two_rank_gloo(tmp_path):
    """N>1 path on CPU: two processes over gloo shard five chromosomes and merge per-round results."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "_dist_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert (tmp_path / "ok").exists() and (tmp_path / "ok_pipe").exists()


def test_columnar_ingest_equals_line_parser(tmp_path):
    """The pandas/numpy ingest must accept, orient and order PETs exactly like the per-line PET class
    (cLoops/io.py:30-59,150-183), including on malformed lines."""
    rng = np.random.default_rng(4)
    lines = []
    for k in range(4000):
        c1 = "chr%d" % rng.integers(1, 4)
        c2 = c1 if rng.random() < 0.8 else "chr%d" % rng.integers(1, 4)
        a = int(rng.integers(0, 100000)); b = int(rng.integers(0, 100000))
        t = [c1, str(a), str(a + int(rng.integers(0, 200))), c2, str(b), str(b + int(rng.integers(0, 200))), "n%d" % k, ".",
             "+-"[rng.integers(0, 2)], "+-"[rng.integers(0, 2)]]
        r = rng.random()
        if r < 0.02: t = t[:int(rng.integers(1, 10))]            # short line
        elif r < 0.04: t[1] = "x12"                               # not an int
        elif r < 0.06: t[4] = " 77"                               # int() accepts, the regex does not: slow path
        elif r < 0.08: t[2] = "+5"                                # signed
        elif r < 0.10: t = ["*", "-1", "-1", "*", "-1", "-1", "n", ".", "+", "-"]
        elif r < 0.12: t += ["extra", "cols"]
        elif r < 0.13: t = [""]
        elif r < 0.14: t[5] = "1_0"                               # python int() accepts underscores
        lines.append("\t".join(t))
    f = tmp_path / "w.bedpe"
    f.write_text("\n".join(lines) + "\n")
    for cs, cut in (([], 0), ({"chr1", "chr3"}, 0), ([], 5000)):
        want = list(io._cis_pets([str(f)], cs, cut, logging.getLogger("t"), True))
        chrom, a, b, opp, n = io._cis_table(str(f), cs, cut)
        got = list(zip(chrom.tolist(), a.tolist(), b.tolist(), opp.tolist()))
        assert got == want
        assert n == len(lines)


def _adversarial_lines(rng, n, crlf=False):
    lines = []
    for k in range(n):
        c1 = "chr%d" % rng.integers(1, 5)
        c2 = c1 if rng.random() < 0.8 else "chr%d" % rng.integers(1, 5)
        a = int(rng.integers(0, 100000)); b = int(rng.integers(0, 100000))
        t = [c1, str(a), str(a + int(rng.integers(0, 200))), c2, str(b), str(b + int(rng.integers(0, 200))), "n%d" % k, ".",
             "+-"[rng.integers(0, 2)], "+-"[rng.integers(0, 2)]]
        r = rng.random()
        if r < 0.02: t = t[:int(rng.integers(1, 10))]            # short line
        elif r < 0.04: t[1] = "x12"                               # not an int
        elif r < 0.06: t[4] = " 77"                               # int() accepts it: an "odd" line, decided by the PET class
        elif r < 0.08: t[2] = "+5"
        elif r < 0.10: t = ["*", "-1", "-1", "*", "-1", "-1", "n", ".", "+", "-"]
        elif r < 0.12: t += ["extra", "cols"]
        elif r < 0.13: t = [""]
        elif r < 0.14: t[5] = "1_0"                               # python 3 int() accepts underscores
        elif r < 0.15: t[1] = "-%d" % a                           # negative coordinate: floor, not truncation (io.py:55)
        elif r < 0.16: t[0] = t[3] = "odd chrom name"
        elif r < 0.17: t[9] = ""                                  # empty trailing field
        elif r < 0.18: t[2] = "1234567890123456789"               # 19 digits: left to python
        elif r < 0.19: t[6] = "*"; t[7] = "-1"                    # the "*" / "-1" test looks at every field (io.py:159)
        lines.append("\t".join(t))
    return lines


def test_native_ingest_equals_line_parser(tmp_path):
    """cloops_bedpe_parse (+ the per-line class for the lines it hands back) accepts, orients and orders PETs exactly like
    the per-line restatement of cLoops/io.py:30-59,150-183: plain and gzip input, several files, CRLF, one block or many,
    one tokenizer or several."""
    rng = np.random.default_rng(11)
    f1 = tmp_path / "a.bedpe"
    f1.write_text("\n".join(_adversarial_lines(rng, 6000)) + "\n")
    f2 = tmp_path / "b.bedpe.gz"
    with gzip.open(f2, "wt") as fh:
        fh.write("\n".join(_adversarial_lines(rng, 5000)))            # no newline at the end
    f3 = tmp_path / "c.bedpe"
    f3.write_bytes(("\r\n".join(_adversarial_lines(rng, 3000)) + "\r\n").encode())
    f4 = tmp_path / "empty.bedpe"
    f4.write_text("")
    log = logging.getLogger("t")
    for fs in ([f1], [f2], [f3], [f4], [f1, f2, f4, f3]):
        fs = [str(f) for f in fs]
        for cs, cut in (([], 0), ({"chr1", "chr3", "odd chrom name"}, 0), ([], 5000)):
            want = list(io._cis_pets(fs, cs, cut, log, True))
            n_lines = io._cis_pets.total
            for threads in (1, 5):
                io.INGEST_THREADS = threads
                try:
                    order, per, total = io._cis_native(fs, cs, cut)
                finally:
                    io.INGEST_THREADS = 0
                assert total == n_lines
                first = {}
                for c, a, b, o in want:
                    first.setdefault(c, []).append((a, b, o))
                assert order == list(first)
                for c in order:
                    a, b, opp, line = per[c]
                    assert list(zip(a.tolist(), b.tolist(), opp.tolist())) == first[c]
                    assert (np.diff(line) > 0).all()
    from cloops_b200._lib import CloopsError
    with pytest.raises(CloopsError, match="cannot open"):      # the reference raises IOError from open() (io.py:150-151)
        io._cis_native([str(tmp_path / "missing.bedpe")], [], 0)
    # a carriage return inside a line: declined, the entry points still answer (line-by-line reader)
    f5 = tmp_path / "cr.bedpe"
    f5.write_bytes(b"chr1\t1\t3\tchr1\t100\t102\tn\t.\t+\t-\rchr1\t5\t7\tchr1\t200\t202\tn\t.\t+\t-\n")
    assert io._cis_native([str(f5)], [], 0) is None
    out = tmp_path / "o"
    os.mkdir(out)
    cfs = io.parseRawBedpe2([str(f5)], str(out), [], 0, log)
    assert io.parseJd(cfs[0])[1].tolist() == [[0, 2, 101], [1, 6, 201]]


def test_native_ingest_many_blocks(tmp_path):
    """A file of many reader blocks (> 4 MB each) with chromosomes interleaved: block stitching keeps file order."""
    rng = np.random.default_rng(12)
    n = 400000
    chrom = rng.integers(1, 24, n)
    a = rng.integers(0, 2 * 10 ** 8, n)
    d = rng.integers(0, 10 ** 6, n)
    with open(tmp_path / "big.bedpe", "w") as fh:
        fh.write("".join("chr%d\t%d\t%d\tchr%d\t%d\t%d\tread_name_padding_%d\t.\t+\t-\n" % (c, x, x + 36, c, x + y, x + y + 36, k)
                         for k, (c, x, y) in enumerate(zip(chrom.tolist(), a.tolist(), d.tolist()))))
    assert os.path.getsize(tmp_path / "big.bedpe") > 5 * (4 << 20)
    order, per, total = io._cis_native([str(tmp_path / "big.bedpe")], [], 0)
    assert total == n
    seen = []
    for c in chrom.tolist():
        if "chr%d" % c not in seen:
            seen.append("chr%d" % c)
    assert order == seen
    for c in range(1, 24):
        m = chrom == c
        A, B, opp, line = per["chr%d" % c]
        assert (A == a[m] + 18).all() and (B == a[m] + d[m] + 18).all() and opp.all()
        assert (line == np.flatnonzero(m)).all()


def test_native_ingest_equals_reference(tmp_path):
    """parseRawBedpe2 / parseRawBedpe + txt2jd of the REFERENCE (cLoops/io.py:62-203, through the shim) against this
    package's entry points on the reference's bundled chr21 example and on an adversarial file: same .jd files, same
    matrices, same opposite-strand distances."""
    if not ref_shim.available():
        pytest.skip("no reference tree")
    example = os.path.join(ref_shim.REF_ROOT, "examples", "GSM1872886_GM12878_CTCF_ChIA-PET_chr21_hg38.bedpe.gz")
    ns = ref_shim.load()
    rng = np.random.default_rng(13)
    adv = tmp_path / "adv.bedpe"
    adv.write_text("\n".join(_adversarial_lines(rng, 5000)) + "\n")
    log = logging.getLogger("t")
    cases = [([str(adv)], [], 0), ([str(adv)], ["chr2", "chr4"], 3000)]
    if os.path.isfile(example):
        cases += [([example], [], 0), ([example, str(adv)], ["chr21", "chr1"], 1000)]
    for k, (fs, cs, cut) in enumerate(cases):
        for dedup in (False, True):
            ro, mo = tmp_path / ("r%d%d" % (k, dedup)), tmp_path / ("m%d%d" % (k, dedup))
            os.mkdir(ro); os.mkdir(mo)
            if dedup:
                want, want_ds = ns.io.parseRawBedpe(fs, str(ro), cs, cut, log)
                got, got_ds = io.parseRawBedpe(fs, str(mo), cs, cut, log)
                assert got_ds == want_ds
            else:
                want = ns.io.parseRawBedpe2(fs, str(ro), cs, cut, log)
                got = io.parseRawBedpe2(fs, str(mo), cs, cut, log)
            want = [ns.io.txt2jd(f) for f in want]
            assert [os.path.basename(f) for f in got] == [os.path.basename(f) for f in want]
            for fw, fg in zip(want, got):
                w, g = joblib.load(fw), joblib.load(fg)
                assert g.dtype == np.int64 and np.array_equal(np.asarray(w, dtype=np.int64), g)


def test_native_ingest_of_the_bundled_example_equals_golden(gold_dir):
    """The reference's only fixture (examples/*.bedpe.gz, copied to oracle/_ref/examples by oracle/make_ref.py so that it
    travels to the GPU box) through cloops_bedpe_parse must give the committed golden coordinates, which oracle/make_golden.py
    produced with the reference's own PET class (cLoops/io.py:30-59)."""
    example = None
    for root in (ref_shim.REF_ROOT, os.path.join(ROOT, "oracle", "_ref")):
        f = os.path.join(root, "examples", "GSM1872886_GM12878_CTCF_ChIA-PET_chr21_hg38.bedpe.gz")
        if os.path.isfile(f):
            example = f
            break
    if example is None:
        pytest.skip("the reference's example file is not available")
    g = np.load(os.path.join(gold_dir, "chr21_pets.npz"))
    order, per, ds = io.readBedpe([example], [], 0, logging.getLogger("t"))
    assert order == ["chr21"] and ds is None
    a, b = per["chr21"]
    assert a.dtype == np.int64 and np.array_equal(a, g["X"]) and np.array_equal(b, g["Y"])
    order, per, ds = io.readBedpe([example], [], 0, logging.getLogger("t"), dedup=True)
    assert len(per["chr21"][0]) == len(g["X"]) and len(ds) > 0          # no duplicate (cA, cB) in the example


def test_facade_argument_contract():
    """Errors that are part of the call surface and need no GPU: empty input (cDBSCAN2 -> {},
    v1/block -> IndexError as in cDBSCAN.py:77 / blockDBSCAN.py:74), malformed mat, non-integer eps."""
    from cloops_b200._lib import CloopsError
    from cloops_b200.blockDBSCAN import blockDBSCAN
    from cloops_b200.cDBSCAN import cDBSCAN as V1
    from cloops_b200.cDBSCAN2 import cDBSCAN as V2
    empty = np.zeros((0, 3), np.int64)
    db = V2(empty, 1000, 5)
    assert db.labels == {} and db.labels_array.shape == (0,)
    for cls in (V1, blockDBSCAN):
        with pytest.raises(IndexError):
            cls(empty, 1000, 5)
    with pytest.raises(CloopsError):
        V2(np.zeros((4, 2), np.int64), 1000, 5)
    with pytest.raises(CloopsError):
        V2(np.zeros((4, 3), np.int64), 0.5, 5)


def test_moment_combination_matches_numpy():
    rng = np.random.default_rng(8)
    for trial in range(20):
        chunks = [rng.integers(1, 3_000_000, int(rng.integers(1, 4000))) for _ in range(int(rng.integers(1, 7)))]
        parts = []
        for c in chunks:
            x = np.log2(c.astype(np.float64))
            parts.append((len(x), float(x.mean()), float(((x - x.mean()) ** 2).sum())))
        allx = np.log2(np.concatenate(chunks).astype(np.float64))
        n, mean, std = ests.combine_moments(parts + [(0, 0.0, 0.0)])
        assert n == len(allx) and abs(mean - allx.mean()) < 1e-12 and abs(std - allx.std()) < 1e-12
    di = rng.integers(5000, 500000, 4001)
    dsv = rng.integers(50, 3000, 6000)
    mom = lambda v: (len(v), float(np.log2(v.astype(float)).mean()), float(((np.log2(v.astype(float)) - np.log2(v.astype(float)).mean()) ** 2).sum()))
    assert ests.cut_from_moments([mom(di[:1000]), mom(di[1000:])], [mom(dsv[:10]), mom(dsv[10:])], np.sort(dsv)) == ests.estIntSelCutFrag(di, dsv)
    assert ests.cut_from_moments([mom(di)], [mom(dsv[:5999])], np.sort(dsv[:5999])) == ests.estIntSelCutFrag(di, dsv[:5999])


def test_scripts_common_windows_and_loop_parsing(gold_dir, tmp_path):
    """scripts/_common: the batched window pairs equal getNearbyPairRegions pair by pair; preDs keeps the
    significant loops of a .loop file in file order and finds iva/ivb by header."""
    from cloops_b200.cModel import getNearbyPairRegions
    from cloops_b200.scripts._common import loop_intervals, nearby_pairs, preDs
    rng = np.random.default_rng(3)
    a0 = rng.integers(0, 5000, 50)
    b0 = a0 + rng.integers(100, 100000, 50)
    iv = np.stack([a0, a0 + rng.integers(0, 3000, 50), b0, b0 + rng.integers(0, 3000, 50)], axis=1)
    got = nearby_pairs(iv)
    for m in range(len(iv)):
        ivas, ivbs = getNearbyPairRegions([int(iv[m, 0]), int(iv[m, 1])], [int(iv[m, 2]), int(iv[m, 3])])
        want = [[a[0], a[1], b[0], b[1]] for a in ivas for b in ivbs]
        assert got[m].tolist() == want
    (tmp_path / "chr21-chr21.jd").write_bytes(b"")
    loop = os.path.join(gold_dir, "chr21_m1.loop")
    rec = preDs(loop, str(tmp_path))
    assert list(rec) == ["chr21"] and rec["chr21"]["f"].endswith("chr21-chr21.jd")
    keys, chroms, ivs = loop_intervals(rec["chr21"]["rs"])
    lines = [l.rstrip("\n").split("\t") for l in open(loop)][1:]
    sig = [l for l in lines if float(l[-1]) >= 1]
    assert keys == [l[0] for l in sig] and len(keys) == 202
    assert ["%s:%d-%d" % (c, r[0], r[1]) for c, r in zip(chroms, ivs)] == [l[10] for l in sig]
    assert preDs(loop, str(tmp_path / "missing")) == {}


def test_deloops_main_keeps_chromosomes_present_in_both_samples(gold_dir, tmp_path, monkeypatch):
    """scripts/deLoops:187-207: only chromosomes with loops AND a .jd file in both samples reach callDeLoops."""
    from cloops_b200.scripts import deLoops
    loop = os.path.join(gold_dir, "chr21_m1.loop")
    lines = open(loop).read().split("\n")
    extra = lines[1].replace("chr21", "chr22")                       # one significant loop on a chromosome only sample a has
    fa = tmp_path / "a.loop"
    fa.write_text("\n".join([lines[0], extra] + lines[1:]))
    da, db = tmp_path / "A", tmp_path / "B"
    da.mkdir()
    db.mkdir()
    for d, chroms in ((da, ("chr21", "chr22")), (db, ("chr21",))):
        for c in chroms:
            (d / ("%s-%s.jd" % (c, c))).write_bytes(b"")
    seen = {}
    monkeypatch.setattr(deLoops, "callDeLoops", lambda ra, rb, prea, preb, dis, cpu: seen.update(ra=ra, rb=rb, pre=(prea, preb), dis=dis))
    monkeypatch.chdir(tmp_path)
    deLoops.main(["-fa", str(fa), "-fb", loop, "-da", str(da), "-db", str(db), "-dis", "7"])
    assert list(seen["ra"]) == ["chr21"] and list(seen["rb"]) == ["chr21"]
    assert seen["pre"] == ("A", "B") and seen["dis"] == 7
    assert len(seen["ra"]["chr21"]["rs"]) == 202
