"""Drop-in for the scoring half of ``cLoops.cModel`` (cLoops/cModel.py:31-331).

The permuted-local-background test needs, per candidate loop, 123 integers (ra, rb, rab, the sizes
of 10+10 shifted windows and their 10x10 joint counts).  They are counted on the GPU by the batched
range-count kernel (``cloops_range_counts``); the statistics on top of them are evaluated on the host
with the SAME numpy/scipy calls the reference makes (cModel.py:114,148-160), so every float in the
loop table is reproduced exactly.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
from scipy.stats import binom, hypergeom, poisson

from . import device
from .io import parseIv, parseJd


class _Axis:
    """One axis of the coverage model: what the reference keeps as ``[sorted coords, {coord: [rows]}]``
    (cModel.py:31-42).  Held as the sorted coordinates plus the argsort permutation."""

    def __init__(self, coords: np.ndarray):
        self.order = np.argsort(coords, kind="stable")
        self.keys = coords[self.order]

    def rows(self, lo, hi) -> np.ndarray:
        a = np.searchsorted(self.keys, lo, side="left")
        b = np.searchsorted(self.keys, hi, side="right")
        return self.order[a:b]


class CoverageModel:
    """Result of getGenomeCoverage.  ``model[0]`` / ``model[1]`` are the X / Y axes (host, lazy, for
    getCounts); ``model.gpu`` is the resident device model the batched kernels use."""

    def __init__(self, X: np.ndarray, Y: np.ndarray):
        self.X = np.ascontiguousarray(X)
        self.Y = np.ascontiguousarray(Y)
        self.N = len(self.X)
        self._axes = [None, None]
        self._gpu = None

    @property
    def gpu(self) -> "device.Coverage":
        if self._gpu is None:
            self._gpu = device.Coverage(device.to_device_i32(self.X, "X"), device.to_device_i32(self.Y, "Y"))
        return self._gpu

    def __getitem__(self, k) -> _Axis:
        if self._axes[k] is None:
            self._axes[k] = _Axis(self.X if k == 0 else self.Y)
        return self._axes[k]

    def __len__(self):
        return 2


#: set by pipe.py: f -> the chromosome object already resident in HBM (key, X, Y, dx, dy), or None
RESIDENT = None
#: the reference reports progress with print() (cModel.py:268-270); bench.py silences it so that its stdout is one JSON line
QUIET = False


def _say(msg):
    if not QUIET:
        print(msg)


def getGenomeCoverage(f, cut=0):
    """cModel.py:45-57: ``(model, N)``; ``(None, 0)`` when fewer than 2 PETs survive ``cut``."""
    ch = RESIDENT(f) if (RESIDENT is not None and cut == 0) else None
    if ch is not None:                                   # pipe(): the chromosome is in HBM already, discut is 0 (pipe.py:284)
        if ch.n < 2:
            return None, 0
        m = CoverageModel(ch.X, ch.Y)
        m._gpu = device.Coverage(ch.dx, ch.dy)
        return m, ch.n
    key, mat = parseJd(f, cut)
    if mat.shape[0] < 2:
        return None, 0
    return CoverageModel(mat[:, 1], mat[:, 2]), mat.shape[0]


def getCounts(iv, model):
    """cModel.py:60-69: set of row indices whose coordinate lies in [iv[0], iv[1]] (inclusive)."""
    return set(model.rows(iv[0], iv[1]).tolist())


def getPETsforRegions(iva, ivb, model):
    """cModel.py:72-80 -> (ra, rb, rab), counted on the GPU."""
    r = model.gpu.region_pets([[iva[0], iva[1], ivb[0], ivb[1]]])[0]
    return int(r[0]), int(r[1]), int(r[2])


def getNearbyPairRegions(iva, ivb, win=5):
    """cModel.py:83-105 with the reference's (py2) integer arithmetic."""
    ca, cb = (iva[0] + iva[1]) // 2, (ivb[0] + ivb[1]) // 2
    sa, sb = (iva[1] - iva[0]) // 2, (ivb[1] - ivb[0]) // 2
    step = (sa + sb) // 2
    shifts = [i for i in range(-win, win + 1) if i != 0]
    ivas = [[max(0, ca + i * step - sa), max(0, ca + i * step + sa)] for i in shifts]
    ivbs = [[max(0, cb + i * step - sb), max(0, cb + i * step + sb)] for i in shifts]
    return ivas, ivbs


def _stats(c: np.ndarray, N: int):
    """cModel.py:113-114,128-161 on the 123 integers of one candidate."""
    ra, rb, rab = int(c[0]), int(c[1]), int(c[2])
    hyp = max([1e-300, hypergeom.sf(rab - 1.0, N, ra, rb)])
    na = c[3:13].astype(np.float64)
    nb = c[13:23].astype(np.int64)
    joint = c[23:123].astype(np.float64).reshape(10, 10)
    rabs = joint.reshape(-1)                                   # zeros stay in the list (:137,143)
    with np.errstate(divide="ignore", invalid="ignore"):
        dens = joint / (na[:, None] * nb[None, :])
    nbps = np.where(joint > 0, dens, 0.0).reshape(-1)
    fdr = len(rabs[rabs > rab]) / float(len(rabs))
    mrabs = float(np.mean(rabs))
    if mrabs > 0:
        es = rab / np.mean(rabs[rabs > 0])
    else:
        es = np.inf
    pop = max([1e-300, poisson.sf(rab - 1.0, mrabs)])
    bp = np.mean(nbps) * ra * rb / N
    nbp = max([1e-300, binom.sf(rab - 1.0, N - rab, bp)])
    return ra, rb, rab, es, fdr, hyp, pop, nbp


_POOL = None


def _chunked(fn, *cols, chunk=2048):
    """fn(*cols) evaluated chunk by chunk on a few host threads: scipy's distribution ufuncs release the GIL and cost tens of
    microseconds per candidate at Hi-C depth (hypergeom.sf with N ~ 10^7); every element is computed exactly as in one call."""
    global _POOL
    n = len(cols[0])
    if n <= chunk:
        return fn(*cols)
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(32, os.cpu_count() or 1)))
    parts = _POOL.map(lambda a: fn(*(c[a:a + chunk] for c in cols)), range(0, n, chunk))
    return np.concatenate(list(parts))


def _stats_batch(counts: np.ndarray, N: int):
    """_stats for K candidates at once -> eight arrays.  Same operations per candidate (the joint counts
    are integers, so their sums are exact in any order; the density mean keeps numpy's row-wise pairwise
    order); poisson/binom/hypergeom are evaluated through their array forms, which return the scalar
    results element by element (checked against the reference's tuples in tests/test_host_logic.py)."""
    c = np.asarray(counts)
    K = c.shape[0]
    ra, rb, rab = (c[:, k].astype(np.int64) for k in range(3))
    na = c[:, 3:13].astype(np.float64)
    nb = c[:, 13:23].astype(np.int64)
    joint = c[:, 23:123].astype(np.float64).reshape(K, 10, 10)
    rabs = joint.reshape(K, 100)
    with np.errstate(divide="ignore", invalid="ignore"):
        dens = joint / (na[:, :, None] * nb[:, None, :])
        nbps = np.where(joint > 0, dens, 0.0).reshape(K, 100)
        fdr = (rabs > rab[:, None]).sum(axis=1) / 100.0
        mrabs = np.mean(rabs, axis=1)
        npos = (rabs > 0).sum(axis=1)
        es = np.where(mrabs > 0, rab / (rabs.sum(axis=1) / np.maximum(npos, 1)), np.inf)
        hyp = np.maximum(1e-300, _chunked(lambda k, a, b: hypergeom.sf(k, N, a, b), rab - 1.0, ra, rb))
        pop = np.maximum(1e-300, _chunked(poisson.sf, rab - 1.0, mrabs))
        bp = np.mean(nbps, axis=1) * ra * rb / N
        nbp = np.maximum(1e-300, _chunked(binom.sf, rab - 1.0, N - rab, bp))
    return ra, rb, rab, es, fdr, hyp, pop, nbp


def _stats_binom(counts: np.ndarray, N: int):
    """The part of _stats_batch removeDup needs (cModel.py:235-258 looks at the binomial p and rab / ra / rb only):
    -> ra, rb, rab, nbp.  Element for element the same operations as _stats_batch."""
    c = np.asarray(counts)
    K = c.shape[0]
    ra, rb, rab = (c[:, k].astype(np.int64) for k in range(3))
    na = c[:, 3:13].astype(np.float64)
    nb = c[:, 13:23].astype(np.int64)
    joint = c[:, 23:123].astype(np.float64).reshape(K, 10, 10)
    with np.errstate(divide="ignore", invalid="ignore"):
        dens = joint / (na[:, :, None] * nb[:, None, :])
        nbps = np.where(joint > 0, dens, 0.0).reshape(K, 100)
        bp = np.mean(nbps, axis=1) * ra * rb / N
        nbp = np.maximum(1e-300, _chunked(binom.sf, rab - 1.0, N - rab, bp))
    return ra, rb, rab, nbp


def getMultiplePsFdr(iva, ivb, model, N, win=5):
    """cModel.py:108-161 -> (ra, rb, rab, es, fdr, hyp, pop, nbp)."""
    if win != 5:
        raise ValueError("the GPU range-count kernel is built for win=5 (the only value the reference uses)")
    c = model.gpu.range_counts([[iva[0], iva[1], ivb[0], ivb[1]]])[0]
    return _stats(c, N)


def getBonPvalues(ps):
    """cModel.py:164-171."""
    ps = np.array(ps)
    ps = ps * len(ps)
    ps[ps > 1.0] = 1.0
    return ps


def checkOneEndOverlap(xa, xb, ya, yb):
    """cModel.py:174-182."""
    if (ya <= xa <= yb) or (ya <= xb <= yb) or (ya <= xa <= xb <= yb):
        return True
    if (xa <= ya <= xb) or (xa <= yb <= xb) or (xa <= ya <= yb <= xb):
        return True
    return False


def checkOverlap(ivai, ivbi, ivaj, ivbj):
    """cModel.py:185-195."""
    if ivai[0] != ivaj[0] or ivbi[0] != ivbj[0]:
        return
    return bool(checkOneEndOverlap(ivai[1], ivai[2], ivaj[1], ivaj[2])
                and checkOneEndOverlap(ivbi[1], ivbi[2], ivbj[1], ivbj[2]))


def _end_overlap(xa, xb, ya, yb):
    """Vectorised checkOneEndOverlap: scalar (xa, xb) against arrays (ya, yb)."""
    return (((ya <= xa) & (xa <= yb)) | ((ya <= xb) & (xb <= yb)) | ((xa <= ya) & (ya <= xb)) | ((xa <= yb) & (yb <= xb)))


def _remove_dup_index(a0, a1, b0, b1, bp, dens, bpcut=1e-5):
    """removeDup on columns (one chromosome): -> indices of the surviving loops in the reference's output order.  The greedy
    grouping and the choice of winners run in host C++ (``cloops_remove_dup``, csrc/removedup.cu); groups whose maximum
    density is shared come back as tie lists and are resolved here with the reference's expression (pandas'
    ``Series.sort_values(ascending=False)`` spelled out: reverse, argsort ascending with numpy's quicksort, reverse)."""
    import ctypes as C
    from . import _lib
    n = len(a0)
    cols = [np.ascontiguousarray(v, dtype=np.int64) for v in (a0, a1, b0, b1)]
    bp = np.ascontiguousarray(bp, dtype=np.float64)
    dens = np.ascontiguousarray(dens, dtype=np.float64)
    keep, tie_start, tie_members = np.empty(max(n, 1), np.int64), np.empty(n + 1, np.int64), np.empty(max(n, 1), np.int64)
    n_keep, n_ties = C.c_int64(0), C.c_int64(0)
    _lib.check(_lib.lib().cloops_remove_dup(*(c.ctypes.data for c in cols), bp.ctypes.data, dens.ctypes.data, n, float(bpcut),
                                            keep.ctypes.data, C.addressof(n_keep), tie_start.ctypes.data, tie_members.ctypes.data,
                                            C.addressof(n_ties)))
    keep = keep[:n_keep.value]
    for g in range(n_ties.value):
        members = tie_members[tie_start[g]:tie_start[g + 1]]
        vals = dens[members]
        order = np.arange(len(vals))[::-1][vals[::-1].argsort(kind="quicksort")][::-1]
        keep[keep == -(g + 1)] = members[int(order[0])]
    return keep


def removeDup(ds, bpcut=1e-5):
    """cModel.py:198-259: greedy grouping of overlapping loops in key order (first member leads its group), then per group
    keep the densest member among those with binomial p <= bpcut.  Same grouping, same winners and same output order as
    the reference; the pair loop runs as a sweep in host C++."""
    keys = list(ds.keys())
    if len(keys) == 0:
        return {}
    ivs = [(parseIv(ds[k]["iva"]), parseIv(ds[k]["ivb"])) for k in keys]
    chroms = set((a[0], b[0]) for a, b in ivs)
    if len(chroms) > 1:                                  # loops of several chromosome pairs: they never overlap (:188-189)
        ids = {c: i for i, c in enumerate(sorted(chroms))}
        off = np.array([ids[(a[0], b[0])] for a, b in ivs], dtype=np.int64) << 40
    else:
        off = np.zeros(len(keys), np.int64)
    a0 = np.array([a[1] for a, _ in ivs], dtype=np.int64) + off
    a1 = np.array([a[2] for a, _ in ivs], dtype=np.int64) + off
    b0 = np.array([b[1] for _, b in ivs], dtype=np.int64)
    b1 = np.array([b[2] for _, b in ivs], dtype=np.int64)
    bp = np.array([ds[k]["binomial_p-value"] for k in keys], dtype=np.float64)
    dens = np.array([float(ds[k]["rab"]) / ds[k]["ra"] / ds[k]["rb"] for k in keys], dtype=np.float64)
    return {keys[i]: ds[keys[i]] for i in _remove_dup_index(a0, a1, b0, b1, bp, dens, bpcut).tolist()}


def countCandidates(f, records, minPts, discut):
    """The GPU half of getIntSig (cModel.py:262-295): coverage model, (ra, rb, rab) of every candidate, the rab / distance
    filters (:284-291) and the 123 permuted-background integers of the survivors.
    -> None (no model: fewer than 2 PETs) or dict(N, names, cand, keep, dist, counts)."""
    _say("Starting estimate significance for %s candidate interactions in %s" % (len(records), f))
    model, N = getGenomeCoverage(f, discut)
    _say("Genomic coverage model built from %s" % f)
    if N == 0:
        _say("No cis-PETs parsed as requiring distance cutoff >%s from %s" % (discut, f))
        return None
    if isinstance(records, np.ndarray):                  # pipe(): int array [K,4] = minX, maxX, minY, maxY
        cand = records.astype(np.int64).reshape(-1, 4)
        names = tuple(os.path.split(f)[1].replace("mem:", "").replace(".jd", "").split("-")[:2])
    else:
        cand = np.array([[r[1], r[2], r[4], r[5]] for r in records], dtype=np.int64).reshape(-1, 4)
        names = (records[0][0], records[0][3]) if len(records) else ("", "")
    cand[:, 0] = np.maximum(cand[:, 0], 0)               # cModel.py:281-282
    cand[:, 2] = np.maximum(cand[:, 2], 0)
    need = max(minPts)
    dist_all = np.abs((cand[:, 2] + cand[:, 3]) / 2.0 - (cand[:, 0] + cand[:, 1]) / 2.0) if len(cand) else np.zeros(0)
    # two launches, as the reference's two steps: (ra, rb, rab) of every candidate (getPETsforRegions, :287), then the 123
    # integers of the permuted background only for the candidates that pass rab >= max(minPts) (:290-295)
    keep = np.zeros(0, np.int64)
    counts = np.zeros((0, 123), np.int32)
    if len(cand):
        rab = model.gpu.region_pets(cand)[:, 2]
        keep = np.flatnonzero((dist_all >= discut) & (rab >= need))
        if len(keep):
            counts = model.gpu.range_counts(cand[keep])
    model.gpu.close()
    return {"N": N, "names": names, "cand": cand, "keep": keep, "dist": dist_all, "counts": counts}


COLUMNS = ["distance", "ra", "rb", "rab", "ES", "FDR", "hypergeometric_p-value", "poisson_p-value", "binomial_p-value", "iva", "ivb"]


def tableFromCounts(c):
    """The host half of getIntSig (cModel.py:295-331): the reference's numpy / scipy statistics on the counted integers,
    key numbering (:280,292), removeDup twice (:318,322), Bonferroni (:327-330).  -> DataFrame or None.  Works on columns;
    the frame it returns equals the reference's ``pd.DataFrame(ds).T`` (object columns, dict insertion order)."""
    if c is None:
        return None
    N, (chrom_a, chrom_b), cand, keep, dist_all, counts = c["N"], c["names"], c["cand"], c["keep"], c["dist"], c["counts"]
    if len(keep) == 0:
        return None
    # removeDup looks at the binomial p and the density only: those are evaluated for every scored candidate, the other
    # statistics (hypergeom.sf costs ~60 us per candidate at N ~ 10^7) for the survivors -- element for element the same calls
    ra, rb, rab, nbp = _stats_binom(counts, N)
    a0, a1, b0, b1 = (cand[keep, k] for k in range(4))
    dens = rab.astype(np.float64) / ra / rb                     # float(rab) / ra / rb (:244)
    idx = np.arange(len(keep))                                  # key number = accepted so far (cModel.py:280,292)
    for _ in range(2):
        idx = idx[_remove_dup_index(a0[idx], a1[idx], b0[idx], b1[idx], nbp[idx], dens[idx])]
        if len(idx) == 0:
            return None
    full = _stats_batch(counts[idx], N)
    assert np.array_equal(full[7], nbp[idx])
    es, fdr, hyp, pop = (np.empty(len(keep), np.float64) for _ in range(4))
    es[idx], fdr[idx], hyp[idx], pop[idx] = full[3], full[4], full[5], full[6]
    obj = lambda v: np.asarray(v)[idx].astype(object)           # python ints / floats, as the reference's dict values
    iva = np.array(["%s:%s-%s" % (chrom_a, x, y) for x, y in zip(a0[idx].tolist(), a1[idx].tolist())], dtype=object)
    ivb = np.array(["%s:%s-%s" % (chrom_a, x, y) for x, y in zip(b0[idx].tolist(), b1[idx].tolist())], dtype=object)
    cols = {"distance": obj(dist_all[keep]), "ra": obj(ra), "rb": obj(rb), "rab": obj(rab), "ES": obj(es), "FDR": obj(fdr),
            "hypergeometric_p-value": obj(hyp), "poisson_p-value": obj(pop), "binomial_p-value": obj(nbp), "iva": iva, "ivb": ivb}
    ds = pd.DataFrame(cols, index=["%s-%s-%s" % (chrom_a, chrom_b, i) for i in idx.tolist()], columns=COLUMNS)
    ds["poisson_p-value_corrected"] = getBonPvalues(ds["poisson_p-value"])
    ds["binomial_p-value_corrected"] = getBonPvalues(ds["binomial_p-value"])
    ds["hypergeometric_p-value_corrected"] = getBonPvalues(ds["hypergeometric_p-value"])
    return ds


def getIntSig(f, records, minPts, discut):
    """cModel.py:262-331.  All candidates of the chromosome are counted in batched kernel launches; filtering (:284-291),
    key numbering (:280,292), de-duplication and Bonferroni follow the reference."""
    return tableFromCounts(countCandidates(f, records, minPts, discut))


def markIntSig(ds, escut=2.0, fdrcut=1e-2, bpcut=1e-3, ppcut=1e-5, hypcut=1e-10):
    """cModel.py:334-363 (ChIA-PET cut-offs)."""
    ok = ((ds["ES"] >= escut) & (ds["FDR"] <= fdrcut) & (ds["hypergeometric_p-value"] <= hypcut)
          & (ds["poisson_p-value"] <= ppcut) & (ds["binomial_p-value"] <= bpcut))
    ds["significant"] = ok.astype(float).values
    return ds


def markIntSigHic(ds, escut=2.0, fdrcut=0.01, bpcut=1e-5, ppcut=1e-5):
    """cModel.py:366-386 (HiChIP / Hi-C cut-offs: FDR strictly below, no hypergeometric test)."""
    ok = ((ds["ES"] >= escut) & (ds["FDR"] < fdrcut) & (ds["poisson_p-value"] <= ppcut) & (ds["binomial_p-value"] <= bpcut))
    ds["significant"] = ok.astype(float).values
    return ds
