"""Second golden set: the reference's own pipe() on a seeded 3-chromosome HiChIP-like synthetic
(multiple eps AND multiple minPts rounds, -hic marks, several chromosomes).
Run: python oracle/make_golden_multi.py   (needs /root/reference)."""
import io
import logging
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloops_b200 import synth  # noqa: E402  (input generator only)
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CHROMS = (("chr1", 45000, 4_000_000), ("chr2", 30000, 3_000_000), ("chrX", 18000, 2_500_000))
EPS, MINPTS = [1000, 2000], [8, 5]


def inputs():
    out = []
    for ci, (name, n, L) in enumerate(CHROMS):
        X, Y = synth.chromosome(n, L, seed=900 + ci, loop_frac=0.3, sigma=600.0, pets_per_loop=30)
        out.append((name, X, Y))
    return out


def write_bedpe(path, data):
    with open(path, "w") as fh:
        for name, X, Y in data:
            for x, y in zip(X.tolist(), Y.tolist()):
                fh.write("%s\t%d\t%d\t%s\t%d\t%d\tp\t.\t+\t-\n" % (name, x, x, name, y, y))


def main():
    ns = ref_shim.load()
    logging.disable(logging.CRITICAL)
    ns.pipe.logger = logging.getLogger("ref")
    tmp = tempfile.mkdtemp(prefix="cloops_gold2_")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        data = inputs()
        write_bedpe("in.bedpe", data)
        cuts = []
        orig = ns.pipe.estIntSelCutFrag

        def spy(di, ds, log=1):
            r = orig(di, ds, log)
            cuts.append((len(di), len(ds), r[0]))
            return r

        ns.pipe.estIntSelCutFrag = spy
        so, se = sys.stdout, sys.stderr
        sys.stdout, sys.stderr = io.StringIO(), io.StringIO()
        try:
            ns.pipe.pipe(["in.bedpe"], "gold", EPS, MINPTS, cpu=1, tmp=0, hic=1)
        finally:
            sys.stdout, sys.stderr = so, se
            ns.pipe.estIntSelCutFrag = orig
        shutil.copy("gold.loop", os.path.join(GOLD, "multi_hic.loop"))
        np.savez_compressed(os.path.join(GOLD, "multi_hic.npz"), cuts=np.array(cuts, np.int64),
                            **{"%s_%s" % (name, ax): arr for name, X, Y in data for ax, arr in (("X", X), ("Y", Y))})
        print("rounds (ndis, ndss, cut):", cuts)
        print("loop lines:", sum(1 for _ in open("gold.loop")))
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


def main_auto_eps():
    """tests/golden/multi_auto_eps.loop: the same input through pipe(fs, out, eps=0, minPts=[6], hic=0),
    i.e. the reference's auto-eps path (parseRawBedpe + estFragSize, cLoops/pipe.py:229-239)."""
    ns = ref_shim.load()
    logging.disable(logging.CRITICAL)
    ns.pipe.logger = logging.getLogger("ref")
    tmp = tempfile.mkdtemp(prefix="cloops_gold3_")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        write_bedpe("in.bedpe", inputs())
        so, se = sys.stdout, sys.stderr
        sys.stdout, sys.stderr = io.StringIO(), io.StringIO()
        try:
            ns.pipe.pipe(["in.bedpe"], "gold", 0, [6], cpu=1, tmp=0, hic=0)
        finally:
            sys.stdout, sys.stderr = so, se
        shutil.copy("gold.loop", os.path.join(GOLD, "multi_auto_eps.loop"))
        print("auto-eps loop lines:", sum(1 for _ in open("gold.loop")))
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
    main_auto_eps()
