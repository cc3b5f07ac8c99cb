"""Generate tests/golden/* by EXECUTING THE REFERENCE (through oracle/ref_shim.py) in the build
container.  Run:  python oracle/make_golden.py      (needs /root/reference; ~2 min)

The reference ships no tests or golden vectors of its own, so every pin for this
path is produced here from the reference's own code on (a) its bundled chr21 example and (b) seeded
synthetic inputs.  The outputs are small .npz / .tsv fixtures that travel to the GPU box.
"""
from __future__ import annotations

import gzip
import io
import logging
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
EXAMPLE = os.path.join(ref_shim.REF_ROOT, "examples", "GSM1872886_GM12878_CTCF_ChIA-PET_chr21_hg38.bedpe.gz")


def labels_array(labels: dict, ids: np.ndarray) -> np.ndarray:
    out = np.full(len(ids), -1, np.int32)
    pos = {int(i): k for k, i in enumerate(ids.tolist())}
    for i, c in labels.items():
        out[pos[int(i)]] = c
    return out


def battery_case(rng):
    """Small seeded inputs biased toward ties, duplicates, crowded cells, tiny eps."""
    n = int(rng.integers(30, 1200))
    span = int(rng.choice([300, 2000, 20000, 200000]))
    eps = int(rng.choice([1, 2, 5, 20, 100, 300, 1000]))
    minPts = int(rng.integers(2, 12))
    q = int(rng.choice([1, 5, 25, 50]))
    if rng.integers(0, 3) == 0:
        X = rng.integers(0, span, n)
        Y = X + rng.integers(0, span, n)
    else:
        k = max(1, n // 30)
        cx = rng.integers(0, span, k)
        cy = cx + rng.integers(0, span, k)
        idx = rng.integers(0, k, n)
        sig = max(1.0, eps * rng.choice([0.3, 1.0, 3.0]))
        X = (cx[idx] + rng.normal(0, sig, n)).astype(np.int64)
        Y = (cy[idx] + rng.normal(0, sig, n)).astype(np.int64)
        nb = n // 3
        X[:nb] = rng.integers(0, span, nb)
        Y[:nb] = X[:nb] + rng.integers(0, span, nb)
    X = np.abs((X // q) * q)
    Y = np.abs((Y // q) * q)
    X, Y = np.minimum(X, Y), np.maximum(X, Y)
    ids = np.arange(n)
    if rng.random() < 0.3:
        ids = np.sort(rng.choice(np.arange(5 * n), n, replace=False))
    return np.stack([ids, X, Y], axis=1).astype(np.int64), eps, minPts


def main():
    ns = ref_shim.load()
    os.makedirs(GOLD, exist_ok=True)

    # ---- config 1 input: chr21 example through the reference's own PET class (io.py:30-59)
    rows = []
    with gzip.open(EXAMPLE, "rt") as fh:
        for line in fh:
            t = line.rstrip("\n").split("\t")
            try:
                p = ns.io.PET(t)
            except Exception:
                continue
            if p.chromA != p.chromB:
                continue
            rows.append((len(rows), p.cA, p.cB))
    mat = np.array(rows, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "chr21_pets.npz"), X=mat[:, 1].astype(np.int32), Y=mat[:, 2].astype(np.int32))

    # ---- labels of the three reference clusterers on chr21
    out = {}
    classes = {"v2": ns.cDBSCAN2, "v1": ns.cDBSCAN, "block": ns.blockDBSCAN}
    for eps in (500, 1000, 2000):
        for name, cls in classes.items():
            out["%s_eps%d_mp5" % (name, eps)] = labels_array(cls(mat, eps, 5).labels, mat[:, 0])
    for eps, mp in ((5000, 20), (2500, 30)):
        for name, cls in classes.items():
            out["%s_eps%d_mp%d" % (name, eps, mp)] = labels_array(cls(mat, eps, mp).labels, mat[:, 0])
    # cut-filtered rounds as pipe -m 1 sees them (non-contiguous ids, pipe.py:59-63)
    d = mat[:, 2] - mat[:, 1]
    for cut, eps in ((4601, 1000), (13532, 2000)):
        sub = mat[d >= cut]
        for name, cls in classes.items():
            out["%s_cut%d_eps%d_mp5" % (name, cut, eps)] = labels_array(cls(sub, eps, 5).labels, sub[:, 0])
    np.savez_compressed(os.path.join(GOLD, "chr21_labels.npz"), **out)

    # ---- seeded battery, all three variants
    rng = np.random.default_rng(20240)
    bat = {}
    ncase = 120
    for c in range(ncase):
        m, eps, mp = battery_case(rng)
        bat["c%d_mat" % c] = m.astype(np.int32)
        bat["c%d_par" % c] = np.array([eps, mp], np.int32)
        for name, cls in classes.items():
            bat["c%d_%s" % (c, name)] = labels_array(cls(m, eps, mp).labels, m[:, 0])
    bat["ncase"] = np.array(ncase)
    np.savez_compressed(os.path.join(GOLD, "battery_labels.npz"), **bat)

    # ---- end-to-end pipe -m 1 on the example: cuts, candidates, scoring ints/tuples, loop table
    logging.disable(logging.CRITICAL)
    ns.pipe.logger = logging.getLogger("ref")
    tmp = tempfile.mkdtemp(prefix="cloops_gold_")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        rounds = []
        orig_run = ns.pipe.runDBSCAN
        orig_est = ns.pipe.estIntSelCutFrag

        def spy_run(fs, eps, minPts, cut=0, cpu=1):
            r = orig_run(fs, eps, minPts, cut, cpu)
            rounds.append({"eps": eps, "minPts": minPts, "cut_in": cut,
                           "nI": sum(len(v["records"]) for v in r[0].values()), "nS": len(r[1]),
                           "ndis": len(r[2]), "ndss": len(r[3]),
                           "records": [rec[1:3] + rec[4:6] for v in r[0].values() for rec in v["records"]]})
            return r

        def spy_est(di, ds, log=1):
            r = orig_est(di, ds, log)
            rounds[-1]["cut_out"] = r[0]
            return r

        captured = {}
        orig_sig = ns.pipe.getIntSig

        def spy_sig(f, records, minPts, discut):
            captured["f"] = f
            captured["records"] = [list(r) for r in records]
            captured["minPts"] = list(minPts)
            tuples = []
            orig_mp = ns.cModel.getMultiplePsFdr

            def spy_mp(iva, ivb, model, N, win=5):
                r = orig_mp(iva, ivb, model, N, win)
                tuples.append((iva[0], iva[1], ivb[0], ivb[1], N) + tuple(r))
                return r

            ns.cModel.getMultiplePsFdr = spy_mp
            try:
                r = orig_sig(f, records, minPts, discut)
            finally:
                ns.cModel.getMultiplePsFdr = orig_mp
            captured["tuples"] = tuples
            return r

        ns.pipe.runDBSCAN = spy_run
        ns.pipe.estIntSelCutFrag = spy_est
        ns.pipe.getIntSig = spy_sig
        stdout = sys.stdout
        sys.stdout = io.StringIO()
        try:
            ns.pipe.pipe([EXAMPLE], "gold", [500, 1000, 2000], [5], cpu=1, tmp=1, hic=0)
        finally:
            sys.stdout = stdout
            ns.pipe.runDBSCAN = orig_run
            ns.pipe.estIntSelCutFrag = orig_est
            ns.pipe.getIntSig = orig_sig
        shutil.copy(os.path.join(tmp, "gold.loop"), os.path.join(GOLD, "chr21_m1.loop"))
        recs = np.array([[r[1], r[2], r[4], r[5]] for r in captured["records"]], dtype=np.int64)
        tup = np.array(captured["tuples"], dtype=np.float64)
        # the 123 integers per scored candidate, by brute force over the reference's own getCounts
        model, N = ns.cModel.getGenomeCoverage(captured["f"], 0)
        ints = []
        for r in recs[:200]:
            iva = [max(0, int(r[0])), int(r[1])]
            ivb = [max(0, int(r[2])), int(r[3])]
            ra, rb, rab = ns.cModel.getPETsforRegions(iva, ivb, model)
            ivas, ivbs = ns.cModel.getNearbyPairRegions(iva, ivb)
            sa = [ns.cModel.getCounts(w, model[0]) | ns.cModel.getCounts(w, model[1]) for w in ivas]
            sb = [ns.cModel.getCounts(w, model[0]) | ns.cModel.getCounts(w, model[1]) for w in ivbs]
            row = [ra, rb, rab] + [len(s) for s in sa] + [len(s) for s in sb]
            row += [len(a & b) for a in sa for b in sb]
            ints.append(row)
        np.savez_compressed(
            os.path.join(GOLD, "chr21_m1_pipe.npz"),
            round_eps=np.array([r["eps"] for r in rounds]), round_minPts=np.array([r["minPts"] for r in rounds]),
            round_cut_in=np.array([r["cut_in"] for r in rounds]), round_cut_out=np.array([r.get("cut_out", -1) for r in rounds]),
            round_nI=np.array([r["nI"] for r in rounds]), round_nS=np.array([r["nS"] for r in rounds]),
            round_ndis=np.array([r["ndis"] for r in rounds]), round_ndss=np.array([r["ndss"] for r in rounds]),
            **{"round%d_records" % i: np.array(r["records"], np.int64).reshape(-1, 4) for i, r in enumerate(rounds)},
            sig_records=recs, sig_N=np.array(N), sig_tuples=tup, sig_ints200=np.array(ints, np.int64))
        print("rounds:", [(r["eps"], r["cut_in"], r.get("cut_out"), r["nI"], r["nS"]) for r in rounds])
        print("candidates:", len(recs), "scored:", len(tup), "N:", N)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
