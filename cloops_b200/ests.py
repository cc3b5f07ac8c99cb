"""Distance cut-off estimation (cLoops/ests.py) -- O(N) numpy on the host; it is the one cross-
chromosome synchronisation point of a clustering round (cLoops/pipe.py:259)."""
from collections import Counter

import numpy as np
import pandas as pd


def estFragSize(ds, top=500):
    """cLoops/ests.py:23-33: median of the 500 most frequent opposite-strand PET distances."""
    s = pd.Series(Counter(ds))
    s.sort_values(inplace=True, ascending=False)
    return int(np.median(s[:top].index))


def estIntSelCutFrag(di, ds, log=1):
    """cLoops/ests.py:36-61 -> (distance cut-off, fragment size); both integers."""
    di = np.abs(np.array(di))
    ds = np.abs(np.array(ds))
    di = di[~np.isnan(di)]
    ds = ds[~np.isnan(ds)]
    di = di[di > 0]
    ds = ds[ds > 0]
    if log:
        di = np.log2(di)
        ds = np.log2(ds)
    cut1 = np.median(ds) + 3 * ds.std()
    cut2 = (ds.mean() * ds.std() + di.mean() * di.std()) / (ds.std() + di.std())
    rcut = int(2 ** min([cut1, cut2]))
    rfrags = int(2 ** np.median(ds))
    return rcut, rfrags


# ---- the same estimate from per-chromosome partial statistics computed on the GPU ------------------
# The reference pools the PET distances of all chromosomes into two Python lists and runs numpy over
# them (cLoops/pipe.py:120-127,259).  At 10^7-10^8 PETs that host pass costs more than the whole GPU
# round, so pipe() reduces on the device instead: per chromosome (n, mean, M2) of log2(distance) for
# the inter- and self-ligation sets, combined with Chan's parallel formula, plus the integer self-
# ligation distances whose order statistics give the median exactly (log2 is monotone).  Floating-point
# summation order already differs between any two runs of the reference (its list order follows dict
# iteration); what is reproduced, and tested against the reference's rounds, is the INTEGER cut-off.


def combine_moments(parts):
    """[(n, mean, M2), ...] -> (n, mean, std) with population std (numpy's default ddof=0)."""
    n, mean, m2 = 0, 0.0, 0.0
    for nb, mb, m2b in parts:
        if nb == 0:
            continue
        if n == 0:
            n, mean, m2 = nb, mb, m2b
            continue
        delta = mb - mean
        tot = n + nb
        m2 = m2 + m2b + delta * delta * n * nb / tot
        mean = mean + delta * nb / tot
        n = tot
    std = float(np.sqrt(m2 / n)) if n else float("nan")
    return n, mean, std


def cut_from_moments(inter_parts, self_parts, self_sorted_distances):
    """(rcut, rfrags) as estIntSelCutFrag; ``self_sorted_distances``: ascending positive self-ligation
    distances (numpy or torch 1-D integer array) of all chromosomes."""
    _, mi, si = combine_moments(inter_parts)
    _, ms, ss = combine_moments(self_parts)
    k = len(self_sorted_distances)
    lo, hi = self_sorted_distances[(k - 1) // 2], self_sorted_distances[k // 2]
    med = (float(np.log2(float(lo))) + float(np.log2(float(hi)))) / 2.0 if lo != hi else float(np.log2(float(lo)))
    cut1 = med + 3 * ss
    cut2 = (ms * ss + mi * si) / (ss + si)
    return int(2 ** min([cut1, cut2])), int(2 ** med)


def cut_from_round(mom, lo, hi, guard=1e-9):
    """estIntSelCutFrag (cLoops/ests.py:36-61) from a round's device-reduced statistics: ``mom`` = n, sum, sum of squares of
    log2|d| of the inter-ligation (0..2) and self-ligation (3..5) distances, ``lo`` / ``hi`` = the two middle order statistics
    of the positive self-ligation distances.  -> integer cut-off, or None when the estimate must be redone from the pooled
    distances: a middle value outside the histogram (1 <= d < 2^20), or 2**cut within ``guard`` (relative) of an integer, where
    the summation order of the moments could flip ``int()``."""
    ni, si, qi, ns, ss, qs = mom[:6]
    if ni <= 0 or ns <= 0 or lo < 1 or hi < 1 or lo >= (1 << 20) or hi >= (1 << 20):
        return None
    mi, ms = si / ni, ss / ns
    sdi = float(np.sqrt(max(qi / ni - mi * mi, 0.0)))
    sds = float(np.sqrt(max(qs / ns - ms * ms, 0.0)))
    med = (float(np.log2(float(lo))) + float(np.log2(float(hi)))) / 2.0 if lo != hi else float(np.log2(float(lo)))
    cut1 = med + 3 * sds
    cut2 = (ms * sds + mi * sdi) / (sds + sdi)
    c = 2 ** min([cut1, cut2])
    if not np.isfinite(c) or abs(c - round(c)) <= guard * max(1.0, c):
        return None
    return int(c)
