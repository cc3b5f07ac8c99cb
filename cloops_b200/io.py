"""File formats either side of the hot path (cLoops/io.py).  BEDPE ingest and the .jd container are
boundary code: same observable behaviour as the reference, written for numpy instead of per-line
Python objects.  Converters that shell out to external tools (jd2washU, jd2hic) are out of scope."""
from __future__ import annotations

import gzip
import os

import joblib
import numpy as np


class PET(object):
    """One BEDPE line (cLoops/io.py:30-59): cis PETs are oriented left <= right by anchor centre and
    get integer centres cA, cB (py2 floor division)."""
    __slots__ = ["chromA", "chromB", "startA", "startB", "endA", "endB", "strandA", "strandB", "cA", "cB", "distance"]

    def __init__(self, d):
        self.chromA, self.startA, self.endA, self.strandA = d[0], int(d[1]), int(d[2]), d[8]
        self.chromB, self.startB, self.endB, self.strandB = d[3], int(d[4]), int(d[5]), d[9]
        if self.chromA != self.chromB:
            self.cA = self.cB = self.distance = None
            return
        if self.startA + self.endA > self.startB + self.endB:
            self.startA, self.startB = self.startB, self.startA
            self.endA, self.endB = self.endB, self.endA
            self.strandA, self.strandB = self.strandB, self.strandA
        self.cA = (self.startA + self.endA) // 2
        self.cB = (self.startB + self.endB) // 2
        self.distance = self.cB - self.cA


def _open(f):
    return gzip.open(f, "rt") if f.endswith(".gz") else open(f)


def _cis_pets(fs, cs, cut, logger, need_strand):
    """Yield (chrom, cA, cB, opposite_strand) for every accepted cis PET, file order
    (filters of cLoops/io.py:158-176)."""
    i = 0
    for f in fs:
        logger.info("Parsing PETs from %s, requiring initial distance cutoff > %s" % (f, cut))
        with _open(f) as fh:
            for line in fh:
                i += 1
                t = line.split("\n")[0].split("\t")
                if "*" in t and "-1" in t:
                    continue
                if len(t) < 6:
                    continue
                try:
                    pet = PET(t)
                except Exception:
                    continue
                if pet.chromA != pet.chromB:
                    continue
                if len(cs) > 0 and pet.chromA not in cs:
                    continue
                if cut > 0 and pet.distance < cut:
                    continue
                yield pet.chromA, pet.cA, pet.cB, (pet.strandA != pet.strandB)
    _cis_pets.total = i


def _write_jd(fout, per_chrom, order):
    cfs = []
    for c in order:
        rows = per_chrom[c]
        mat = np.empty((len(rows) // 2, 3), dtype=np.int64)
        mat[:, 0] = np.arange(mat.shape[0])
        mat[:, 1] = rows[0::2]
        mat[:, 2] = rows[1::2]
        f = os.path.join(fout, "%s-%s.jd" % (c, c))
        joblib.dump(mat, f)
        cfs.append(f)
    return cfs


def parseRawBedpe2(fs, fout, cs, cut, logger):
    """cLoops/io.py:132-189 + txt2jd (:192-203) in one step: per-chromosome ``[id, cA, cB]`` int64
    matrices (id restarts at 0 per chromosome, rows in file order) written straight to ``.jd``.
    Returns the list of .jd paths in order of first appearance."""
    per, order, j = {}, [], 0
    for c, a, b, _ in _cis_pets(fs, cs, cut, logger, False):
        if c not in per:
            per[c] = []
            order.append(c)
        per[c].append(a)
        per[c].append(b)
        j += 1
    logger.info("Totaly %s PETs from %s, in which %s cis PETs" % (getattr(_cis_pets, "total", 0), ",".join(fs), j))
    return _write_jd(fout, per, order)


def parseRawBedpe(fs, fout, cs, cut, logger):
    """cLoops/io.py:62-129: as parseRawBedpe2 but drops duplicate (cA, cB) per chromosome and collects
    the distances of opposite-strand PETs (input of estFragSize when eps is auto-estimated)."""
    per, order, seen, ds, j = {}, [], {}, [], 0
    for c, a, b, opp in _cis_pets(fs, cs, cut, logger, True):
        if c not in per:
            per[c] = []
            seen[c] = set()
            order.append(c)
        if (a, b) in seen[c]:
            continue
        seen[c].add((a, b))
        per[c].append(a)
        per[c].append(b)
        j += 1
        if opp:
            ds.append(b - a)
    logger.info("Totaly %s PETs from %s, in which %s cis PETs" % (getattr(_cis_pets, "total", 0), ",".join(fs), j))
    return _write_jd(fout, per, order), ds


def txt2jd(f):
    """cLoops/io.py:192-203: tab-separated ``id x y`` text -> joblib .jd (int64 [N,3]); removes the text file."""
    data = np.loadtxt(f, dtype=np.int64, delimiter="\t", ndmin=2)
    fo = f.replace(".txt", ".jd")
    joblib.dump(data, fo)
    os.remove(f)
    return fo


def parseJd(f, cut=0):
    """cLoops/io.py:206-217: ``((chrA, chrB), mat)``; rows with Y-X < cut dropped when cut > 0."""
    key = tuple(os.path.split(f)[1].replace(".jd", "").split("-"))
    mat = joblib.load(f)
    if cut > 0:
        mat = mat[(mat[:, 2] - mat[:, 1]) >= cut, :]
    return key, mat


def parseIv(iv):
    """cLoops/io.py:242-248: "chr:start-end" -> [chr, start, end]."""
    c, rest = iv.split(":")[0], iv.split(":")[1]
    return [c, int(rest.split("-")[0]), int(rest.split("-")[1])]


def _loop_rows(fin, significant):
    with open(fin) as fh:
        header = fh.readline().rstrip("\n").split("\t")
        col = {name: k for k, name in enumerate(header)}
        for line in fh:
            t = line.rstrip("\n").split("\t")
            if significant and float(t[col["significant"]]) < 1:
                continue
            yield col, t


def loops2washU(fin, fout, logger, significant=1):
    """cLoops/io.py:220-239: washU long-range track, one line per loop ``iva ivb 1``.  Columns are
    looked up by NAME (the reference indexes by position, which only matches pandas<0.25 ordering)."""
    logger.info("Converting %s to washU long range interaction track." % fin)
    with open(fout, "w") as f:
        for col, t in _loop_rows(fin, significant):
            f.write("\t".join([t[col["iva"]], t[col["ivb"]], "1"]) + "\n")
    logger.info("Converting %s to washU long range interaction track finished." % fin)


def loops2juice(fin, fout, logger, significant=1):
    """cLoops/io.py:251-289: Juicebox 2D annotation; p-values as -log10."""
    logger.info("Converting %s to Juicebox 2D annotation feature." % fin)
    head = ["chromosome1", "x1", "x2", "chromosome2", "y1", "y2", "color", "observed", "loopId", "FDR",
            "EnrichmentScore", "distance", "-log10(binomal_p-value)", "-log10(poisson_p-value)",
            "-log10(hypergeometric_p-value)"]
    with open(fout, "w") as f:
        f.write("\t".join(head) + "\n")
        for col, t in _loop_rows(fin, significant):
            iva, ivb = parseIv(t[col["iva"]]), parseIv(t[col["ivb"]])
            try:
                row = [iva[0], iva[1], iva[2], ivb[0], ivb[1], ivb[2], '"0,255,255"', t[col["rab"]], t[col["loopId"]],
                       t[col["FDR"]], t[col["ES"]], t[col["distance"]], -np.log10(float(t[col["binomial_p-value"]])),
                       -np.log10(float(t[col["poisson_p-value"]])), -np.log10(float(t[col["hypergeometric_p-value"]]))]
            except Exception:
                continue
            f.write("\t".join(map(str, row)) + "\n")
    logger.info("Converting %s to Juicebox 2D annotation feature finished." % fin)
