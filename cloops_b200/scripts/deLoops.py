"""Drop-in for scripts/deLoops: differentially enriched loops between two samples (each loop of one sample
tested against the other sample's PETs, Poisson, Bonferroni).

Reference behaviour that is reproduced on purpose: scripts/deLoops:73-99 ``getPermutatedBg`` hands the two-axis
model to ``cModel.getCounts`` (which expects ONE axis); the exception is swallowed by a bare ``except:
continue`` for every window, so the permuted background is 0.0 for every loop and
``lam = max(1, rabc + 1) * Nt / Nc`` (:113-114).  The golden files come from the reference run as it is
(oracle/make_golden_scripts.py), so this port returns the same numbers.

  python -m cloops_b200.scripts.deLoops -fa a.loop -fb b.loop -da A -db B [-p cpu] [-c chroms] [-dis cut]
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import pandas as pd
from scipy.stats import poisson

from ..cModel import getBonPvalues, getGenomeCoverage
from ..utils import getLogger
from ._common import loop_intervals, preDs

logger = None


def deloopHelp(argv=None):
    """Same flags as cLoops/utils.py:207-262 (deloopHelp)."""
    parser = argparse.ArgumentParser(description="Differentially enriched loops between two cLoops runs: "
                                                 "deLoops -fa a.loop -fb b.loop -da A -db B")
    jd = "directory with the %s sample's .jd files (written by cLoops -s 1)"
    parser.add_argument("-fa", dest="fa", required=True, type=str, help=".loop file of sample a; loops marked significant (last column 1) are used")
    parser.add_argument("-fb", dest="fb", required=True, type=str, help=".loop file of sample b")
    parser.add_argument("-da", dest="da", required=True, type=str, help=jd % "a")
    parser.add_argument("-db", dest="db", required=True, type=str, help=jd % "b")
    parser.add_argument("-p", dest="cpu", default=1, type=int, help="accepted for compatibility; chromosomes run on the GPU in turn")
    parser.add_argument("-c", dest="chroms", default="", type=str, help="restrict to these chromosomes, e.g. chr1,chr2")
    parser.add_argument("-dis", dest="dis", default=0, type=int, help="drop PETs closer than this distance before counting (default 0)")
    return parser.parse_args(argv)


def getPermutatedBg(ivas, ivbs, model):
    """scripts/deLoops:73-99 as it behaves: 0.0 (see the module docstring)."""
    return 0.0


def estSigTvsC(rs, modelt, Nt, modelc, Nc, pre):
    """scripts/deLoops:121-152 for all loops of one chromosome at once."""
    keys, chroms, iv = loop_intervals(rs)
    if len(keys) == 0:
        return None
    normratio = float(Nt) / float(Nc)
    rabt = modelt.gpu.region_pets(iv)[:, 2].astype(np.int64)
    rabc = modelc.gpu.region_pets(iv)[:, 2].astype(np.int64)
    mrabc = 0.0
    lam = np.maximum((mrabc + 1.0) * normratio, (rabc + 1.0) * normratio)           # :113-114
    pop = np.maximum(poisson.sf(rabt - 1.0, lam), 1e-300)
    fc = rabt / lam
    ds = {}
    for k in range(len(keys)):
        ds[keys[k]] = {
            "iva": "%s:%s-%s" % (chroms[k], iv[k, 0], iv[k, 1]),
            "ivb": "%s:%s-%s" % (chroms[k], iv[k, 2], iv[k, 3]),
            "poisson_p-value": float(pop[k]),
            "FoldEnrichment": float(fc[k]),
        }
    ds = pd.DataFrame(ds).T
    ds["poisson_p-value_corrected"] = getBonPvalues(ds["poisson_p-value"])
    return ds


def estSigOneChr(rst, bedpet, rsc, bedpec, pre, dis=0):
    """scripts/deLoops:155-171."""
    modelt, Nt = getGenomeCoverage(bedpet, dis)
    modelc, Nc = getGenomeCoverage(bedpec, dis)
    dst = estSigTvsC(rst, modelt, Nt, modelc, Nc, pre)
    dsc = estSigTvsC(rsc, modelc, Nc, modelt, Nt, pre)
    return dst, dsc


def callDeLoops(ra, rb, prea, preb, dis=0, cpu=1):
    ds = [estSigOneChr(ra[key]["rs"], ra[key]["f"], rb[key]["rs"], rb[key]["f"], key, dis) for key in ra.keys()]
    dsa = [t[0] for t in ds if t[0] is not None]
    dsb = [t[1] for t in ds if t[1] is not None]
    dsa, dsb = pd.concat(dsa), pd.concat(dsb)
    dsa.to_csv(prea + ".deloop", sep="\t", index_label="loopId")
    dsb.to_csv(preb + ".deloop", sep="\t", index_label="loopId")


def main(argv=None):
    global logger
    logger = getLogger(os.path.join(os.getcwd(), "deLoops.log"))
    op = deloopHelp(argv)
    chroms = set(op.chroms.split(",")) if op.chroms else []
    ra = preDs(op.fa, op.da, chroms, logger=logger)
    rb = preDs(op.fb, op.db, chroms, logger=logger)
    both = set(ra) & set(rb)                              # scripts/deLoops:196-205: chromosomes present in both samples
    for recs, loops, jds in ((ra, op.fa, op.da), (rb, op.fa, op.da)):
        for key in [k for k in recs if k not in both]:
            del recs[key]
            logger.info("No match of %s in %s or %s" % (key, loops, jds))
    callDeLoops(ra, rb, os.path.split(op.da)[1], os.path.split(op.db)[1], op.dis, op.cpu)


if __name__ == "__main__":
    main()
