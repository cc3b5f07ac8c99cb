"""Drop-in for scripts/quantifyLoops.py: ra, rb, rab, enrichment score and Poisson p-value of every
significant loop of a .loop file against the PETs of a directory of .jd files.

Per chromosome the reference walks loops x 100 shifted window pairs through getPETsforRegions
(scripts/quantifyLoops.py:131-181), i.e. 4 x 101 Python set constructions per loop; here all 101 x loops
(iva, ivb) pairs of a chromosome go through ONE cloops_region_pets launch, and the statistics are the same
scipy calls on the same integers.

  python -m cloops_b200.scripts.quantifyLoops -f a.loop -d A -o fout [-p cpu] [-c chr1,chr2] [-dis cut]
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import pandas as pd
from scipy.stats import poisson

from ..cModel import getGenomeCoverage
from ..utils import getLogger
from ._common import loop_intervals, nearby_pairs, preDs

logger = None


def help(argv=None):
    """Same flags as scripts/quantifyLoops.py:29-89."""
    parser = argparse.ArgumentParser(description="Quantify the significant loops of a cLoops run against a set of PETs: "
                                                 "quantifyLoops -f a.loop -d A -o fout")
    parser.add_argument("-f", dest="f", required=True, type=str, help=".loop file; loops marked significant (last column 1) are used")
    parser.add_argument("-d", dest="d", required=True, type=str, help="directory with the sample's .jd files (written by cLoops -s 1)")
    parser.add_argument("-o", dest="output", required=True, type=str, help="prefix of the output table")
    parser.add_argument("-p", dest="cpu", default=1, type=int, help="accepted for compatibility; chromosomes run on the GPU in turn")
    parser.add_argument("-c", dest="chroms", default="", type=str, help="restrict to these chromosomes, e.g. chr1,chr2")
    parser.add_argument("-dis", dest="dis", default=0, type=int, help="drop PETs closer than this distance before counting (default 0)")
    return parser.parse_args(argv)


def estSigOneChr(rs, jdf, pre, dis=0, win=5):
    """scripts/quantifyLoops.py:145-184 -> DataFrame (loops x [iva, ivb, ra, rb, rab, ES, poisson_p-value]) or None."""
    if logger is not None:
        logger.info("Building genomic coverage model for %s" % jdf)
    model, N = getGenomeCoverage(jdf, dis)
    keys, chroms, iv = loop_intervals(rs)
    if len(keys) == 0:
        return None
    m = len(keys)
    cand = np.concatenate([iv[:, None, :], nearby_pairs(iv, win)], axis=1)          # [m, 1 + 100, 4]
    counts = model.gpu.region_pets(cand.reshape(-1, 4)).reshape(m, -1, 3).astype(np.int64)
    ra, rb, rab = counts[:, 0, 0], counts[:, 0, 1], counts[:, 0, 2]
    mrab = counts[:, 1:, 2].sum(axis=1) / float(counts.shape[1] - 1)                # mean of the 100 permuted rab (:131-142)
    with np.errstate(divide="ignore", invalid="ignore"):
        es = np.where(mrab > 0, rab / mrab, 100.0)
    pop = np.maximum(1e-300, poisson.sf(rab - 1.0, mrab))
    ds = {}
    for k in range(m):
        ds[keys[k]] = {
            "iva": "%s:%s-%s" % (chroms[k], iv[k, 0], iv[k, 1]),
            "ivb": "%s:%s-%s" % (chroms[k], iv[k, 2], iv[k, 3]),
            "ra": int(ra[k]), "rb": int(rb[k]), "rab": int(rab[k]),
            "ES": float(es[k]) if mrab[k] > 0 else 100,
            "poisson_p-value": float(pop[k]),
        }
    return pd.DataFrame(ds).T


def quantifyLoops(ra, prea, dis=0, cpu=1):
    ds = [estSigOneChr(ra[key]["rs"], ra[key]["f"], key, dis) for key in ra.keys()]
    ds = [d for d in ds if d is not None]
    ds = pd.concat(ds)
    ds.to_csv(prea + "_quantLoops.txt", sep="\t", index_label="loopId")


def main(argv=None):
    global logger
    logger = getLogger(os.path.join(os.getcwd(), "quantifyLoops.log"))
    op = help(argv)
    chroms = [] if op.chroms == "" else set(op.chroms.split(","))
    ra = preDs(op.f, op.d, chroms, logger=logger)
    quantifyLoops(ra, op.output, op.dis, op.cpu)


if __name__ == "__main__":
    main()
