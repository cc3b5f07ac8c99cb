"""sha256 digests of the C oracle's outputs on the FULL-SIZE inputs BASELINE.json names, computed on the CPU of the
build container (no GPU, no reference needed -- oracle/coracle.c is pinned to the reference by tests/test_coracle.py):

    python oracle/make_digests.py            (~3 min)  -> tests/golden/fullsize_digests.json

tests/test_gpu_fullsize.py compares the CUDA path's labels / records / range counts with the live C oracle AND with
these committed digests (which also pin the synthetic generator and the oracle against drift)."""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from cloops_b200 import synth  # noqa: E402
from oracle import coracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")

#: name -> (input factory, [(variant, eps, minPts, cut)])
CASES = {
    "config2_10M": (lambda: synth.config2(10_000_000), [(2, 1000, 5, 0), (1, 1000, 5, 0), (3, 1000, 5, 0)]),
    "config4_chr1_16M": (lambda: synth.genome_chrom(200_000_000, 4, 0)[1:],
                         [(2, 5000, 50, 0), (2, 5000, 20, 0), (2, 10000, 20, 0), (2, 7500, 30, 6000), (1, 5000, 20, 0), (3, 10000, 50, 0)]),
    "config4_chr21_3M": (lambda: synth.genome_chrom(200_000_000, 4, 20)[1:], [(2, 5000, 50, 0), (2, 10000, 20, 0), (1, 7500, 40, 0), (3, 5000, 20, 0)]),
    "config3_chr1_8M": (lambda: synth.genome_chrom(100_000_000, 3, 0)[1:], [(2, 2500, 30, 0), (2, 10000, 20, 0)]),
}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case_digest(X, Y, variant, eps, minPts, cut, score=False):
    X64, Y64 = X.astype(np.int64), Y.astype(np.int64)
    m = (Y64 - X64) >= cut
    lab = np.full(len(X), -1, np.int32)
    lab[m], info = coracle.dbscan(X64[m], Y64[m], eps, minPts, variant, return_info=True)
    bbox, size, kind = coracle.cluster_records(X64, Y64, lab)
    bbox[size == 0] = 0                      # ids without members (v1 keeps the gaps of deleted clusters, cDBSCAN.py:149-152)
    d = {"labels": sha(lab), "n_labelled": int((lab >= 0).sum()), "n_clusters": int(info["clusters"]), "n_dead": int(info["dead"]),
         "bbox": sha(bbox.astype(np.int32)), "kind": sha(kind), "n_inter": int((kind == 1).sum())}
    if score:
        cand = bbox[kind == 1].copy()
        cand[:, 0] = np.maximum(cand[:, 0], 0)
        cand[:, 2] = np.maximum(cand[:, 2], 0)
        d["counts"] = sha(coracle.range_counts(X64, Y64, cand).astype(np.int32))
    return d


def main():
    out = {}
    for name, (make, runs) in CASES.items():
        X, Y = make()
        out[name] = {"n": int(len(X)), "input": sha(np.stack([X, Y]))}
        for k, (variant, eps, minPts, cut) in enumerate(runs):
            t0 = time.time()
            key = "v%d_eps%d_mp%d_cut%d" % (variant, eps, minPts, cut)
            out[name][key] = case_digest(X, Y, variant, eps, minPts, cut, score=(k == 0))
            print(name, key, out[name][key]["n_clusters"], out[name][key]["n_inter"], "%.1f s" % (time.time() - t0), flush=True)
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
