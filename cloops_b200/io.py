"""File formats either side of the hot path (cLoops/io.py).  BEDPE ingest and the .jd container are
boundary code: same observable behaviour as the reference.  The text is read by the native reader of
libcloops_b200 (``cloops_bedpe_parse``: one inflate/read thread + a pool of tokenizers, csrc/ingest.cu);
lines it does not decide (coordinates that are not plain decimal integers) go through the per-line PET class
below, i.e. through the reference's own expression.  Converters that shell out to external tools (jd2washU,
jd2hic) are out of scope."""
from __future__ import annotations

import ctypes as C
import gzip
import os

import joblib
import numpy as np

from . import _lib


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


def _write_jd(fout, chrom_arrays, order):
    cfs = []
    for c in order:
        a, b = chrom_arrays[c][:2]
        mat = np.empty((len(a), 3), dtype=np.int64)
        mat[:, 0] = np.arange(len(a))
        mat[:, 1] = a
        mat[:, 2] = b
        f = os.path.join(fout, "%s-%s.jd" % (c, c))
        joblib.dump(mat, f)
        cfs.append(f)
    return cfs


_INT_RE = r"^[+-]?[0-9]+$"


def _cis_table(f, cs, cut):
    """Columnar ingest of one BEDPE file: the C tokenizer of pandas splits the lines, numpy does the PET
    arithmetic (cLoops/io.py:47-58).  Lines the reference would have to think about (non-plain integers,
    fewer than 10 fields) are re-parsed one by one through the PET class so that the accept/reject
    decision is exactly the reference's; everything is returned in file order.
    -> (chrom [object], cA [int64], cB [int64], opposite_strand [bool], n_lines)"""
    import pandas as pd
    width = 24
    try:
        tab = pd.read_csv(f, sep="\t", header=None, names=list(range(width)), dtype=str, quoting=3, na_filter=False,
                          keep_default_na=False, engine="c", skip_blank_lines=False, compression="infer")
    except Exception:
        tab = None
    if tab is None or len(tab) == 0:
        rows = list(_cis_pets([f], cs, cut, _NullLog(), True))
        return (np.array([r[0] for r in rows], dtype=object), np.array([r[1] for r in rows], dtype=np.int64),
                np.array([r[2] for r in rows], dtype=np.int64), np.array([r[3] for r in rows], dtype=bool), getattr(_cis_pets, "total", 0))
    n = len(tab)
    cols = [tab[k] for k in range(width)]
    filled = np.column_stack([c.to_numpy(dtype=object) != "" for c in cols])          # trailing pads are ""
    nfields = np.where(filled.any(axis=1), width - np.argmax(filled[:, ::-1], axis=1), 1)
    vals = np.column_stack([c.to_numpy(dtype=object) for c in cols])
    star = ((vals == "*") & filled).any(axis=1) & (vals == "-1").any(axis=1)           # io.py:159
    plain = np.ones(n, dtype=bool)
    for k in (1, 2, 4, 5):
        plain &= cols[k].str.match(_INT_RE).to_numpy()
    # an overlong line or a line ending in empty fields is left to the slow path
    fast = plain & (nfields >= 10) & (nfields < width) & ~star
    slow = ~fast & ~star & (nfields >= 6)
    cA = np.zeros(n, dtype=np.int64)
    cB = np.zeros(n, dtype=np.int64)
    keep = np.zeros(n, dtype=bool)
    opp = np.zeros(n, dtype=bool)
    chrom = cols[0].to_numpy(dtype=object)
    if fast.any():
        idx = np.flatnonzero(fast)
        sA = cols[1].to_numpy(dtype=object)[idx].astype(np.int64)
        eA = cols[2].to_numpy(dtype=object)[idx].astype(np.int64)
        sB = cols[4].to_numpy(dtype=object)[idx].astype(np.int64)
        eB = cols[5].to_numpy(dtype=object)[idx].astype(np.int64)
        cis = chrom[idx] == cols[3].to_numpy(dtype=object)[idx]
        swap = (sA + eA) > (sB + eB)                                                   # io.py:51-54
        a = np.where(swap, (sB + eB) // 2, (sA + eA) // 2)
        b = np.where(swap, (sA + eA) // 2, (sB + eB) // 2)
        ok = cis.copy()
        if len(cs) > 0:
            ok &= np.isin(chrom[idx], list(cs))
        if cut > 0:
            ok &= (b - a) >= cut
        cA[idx], cB[idx], keep[idx] = a, b, ok
        opp[idx] = cols[8].to_numpy(dtype=object)[idx] != cols[9].to_numpy(dtype=object)[idx]
    raw = None
    slow_idx = np.flatnonzero(slow).tolist()
    if slow_idx:
        # pandas pads short rows with "", so a row's true field count is only known up to its trailing empty fields: the
        # slow path splits the raw text of those lines itself, exactly as the reference does (io.py:154)
        with _open(f) as fh:
            raw = fh.read().split("\n")
        if len(raw) < n or (len(raw) > n and any(raw[n:])):       # line structure differs from the tokenizer's: per-line parser
            raw = None
    if slow_idx and raw is None:
        rows = list(_cis_pets([f], cs, cut, _NullLog(), True))
        return (np.array([r[0] for r in rows], dtype=object), np.array([r[1] for r in rows], dtype=np.int64),
                np.array([r[2] for r in rows], dtype=np.int64), np.array([r[3] for r in rows], dtype=bool), getattr(_cis_pets, "total", 0))
    for k in slow_idx:
        t = raw[k].split("\t")
        if ("*" in t and "-1" in t) or len(t) < 6:
            continue
        try:
            pet = PET(t)
        except Exception:
            continue
        if pet.chromA != pet.chromB or (len(cs) > 0 and pet.chromA not in cs) or (cut > 0 and pet.distance < cut):
            continue
        cA[k], cB[k], keep[k], opp[k] = pet.cA, pet.cB, True, pet.strandA != pet.strandB
    return chrom[keep], cA[keep], cB[keep], opp[keep], n


class _NullLog:
    def info(self, *a, **k):
        pass


INGEST_THREADS = 0          # 0: one tokenizer per host core


def _accept_odd(text, cs, cut):
    """One line the native reader handed back: the reference's per-line decision (io.py:154-176)."""
    t = text.split("\t")
    if ("*" in t and "-1" in t) or len(t) < 6:
        return None
    try:
        pet = PET(t)
    except Exception:
        return None
    if pet.chromA != pet.chromB or (len(cs) > 0 and pet.chromA not in cs) or (cut > 0 and pet.distance < cut):
        return None
    return pet.chromA, pet.cA, pet.cB, pet.strandA != pet.strandB


def _cis_native(fs, cs, cut):
    """All files through ``cloops_bedpe_parse``.  -> (order, {chrom: (cA, cB, opposite, line_no)}, n_lines);
    chromosomes in order of first appearance, PETs in file order (io.py:177-185).  None when a file holds a carriage
    return outside "\\r\\n" (python's text mode would split the line there): the caller reads such input line by line."""
    L = _lib.lib()
    paths = (C.c_char_p * max(len(fs), 1))(*[os.fsencode(f) for f in fs])
    names = sorted(cs)
    wanted = (C.c_char_p * max(len(names), 1))(*[c.encode() for c in names])
    h = C.c_void_p()
    _lib.check(L.cloops_bedpe_parse(paths, len(fs), wanted, len(names), int(cut), INGEST_THREADS, C.byref(h)))
    try:
        if L.cloops_bedpe_bare_cr(h) > 0:
            return None
        order, per = [], {}
        ln, npets, where = C.c_int64(), C.c_int64(), C.c_int64()
        for k in range(L.cloops_bedpe_n_chroms(h)):
            ptr = L.cloops_bedpe_chrom(h, k, C.byref(ln), C.byref(npets))
            name = C.string_at(ptr, ln.value).decode()
            a = np.empty(npets.value, np.int64)
            b = np.empty(npets.value, np.int64)
            opp = np.empty(npets.value, np.uint8)
            line = np.empty(npets.value, np.int64)
            _lib.check(L.cloops_bedpe_fetch(h, k, a.ctypes.data, b.ctypes.data, opp.ctypes.data, line.ctypes.data))
            order.append(name)
            per[name] = (a, b, opp.view(bool), line)
        extra = {}
        for k in range(L.cloops_bedpe_n_odd(h)):
            ptr = L.cloops_bedpe_odd(h, k, C.byref(where), C.byref(ln))
            got = _accept_odd(C.string_at(ptr, ln.value).decode(), cs, cut)
            if got is not None:
                extra.setdefault(got[0], []).append((where.value,) + got[1:])
        if extra:                                   # merge by line number; chromosome order = first accepted line
            none = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, bool), np.zeros(0, np.int64))
            for name, rows in extra.items():
                a, b, opp, line = per.get(name, none)
                line = np.concatenate([line, np.array([r[0] for r in rows], np.int64)])
                a = np.concatenate([a, np.array([r[1] for r in rows], np.int64)])
                b = np.concatenate([b, np.array([r[2] for r in rows], np.int64)])
                opp = np.concatenate([opp, np.array([r[3] for r in rows], bool)])
                o = np.argsort(line, kind="stable")
                per[name] = (a[o], b[o], opp[o], line[o])
            order = sorted(per, key=lambda c: per[c][3][0])
        return order, per, int(L.cloops_bedpe_lines(h))
    finally:
        L.cloops_bedpe_free(h)


def _cis_all(fs, cs, cut, logger):
    """-> (order, {chrom: (cA, cB, opposite, seq)}, n_lines) over all files, seq increasing in file order: the native
    reader, or -- for input with bare carriage returns -- the tokenizer of pandas with the per-line class behind it."""
    for f in fs:
        logger.info("Parsing PETs from %s, requiring initial distance cutoff > %s" % (f, cut))
    got = _cis_native(fs, cs, cut)
    if got is not None:
        return got
    parts, total = [], 0
    for f in fs:
        chrom, a, b, opp, n = _cis_table(f, cs, cut)
        parts.append((chrom, a, b, opp))
        total += n
    chrom = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, dtype=object)
    a = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int64)
    b = np.concatenate([p[2] for p in parts]) if parts else np.zeros(0, np.int64)
    opp = np.concatenate([p[3] for p in parts]) if parts else np.zeros(0, bool)
    seq = np.arange(len(a), dtype=np.int64)
    import pandas as pd
    codes, uniques = pd.factorize(chrom, sort=False)
    per, order = {}, []
    for k, name in enumerate(uniques.tolist()):
        m = codes == k
        per[name] = (a[m], b[m], opp[m], seq[m])
        order.append(name)
    return order, per, total


def readBedpe(fs, cs, cut, logger, dedup=False):
    """The PETs parseRawBedpe2 (dedup False, cLoops/io.py:132-189) or parseRawBedpe (dedup True, :62-129) would write, kept
    in memory: -> (order, {chrom: (cA, cB)}, ds).  ``order``: chromosomes by first appearance; cA, cB int64 in file order
    (the row id of the reference's text file is the position); ``ds``: with dedup, the distances of the kept PETs whose
    strands differ, in file order over all chromosomes (:126-127), else None.  dedup drops a PET whose (cA, cB) was seen
    before on its chromosome (:114-115)."""
    order, per, total = _cis_all(fs, cs, cut, logger)
    kept, ds, seqs = {}, [], []
    for c in order:
        a, b, opp, seq = per[c]
        if dedup:
            import pandas as pd
            first = ~pd.DataFrame({"a": a, "b": b}).duplicated(keep="first").to_numpy() if len(a) else np.zeros(0, bool)
            kept[c] = (a[first], b[first])
            ds.append((b - a)[first & opp])
            seqs.append(seq[first & opp])
        else:
            kept[c] = (a, b)
    logger.info("Totaly %s PETs from %s, in which %s cis PETs" % (total, ",".join(fs), sum(len(kept[c][0]) for c in order)))
    if not dedup:
        return order, kept, None
    if ds:
        ds = np.concatenate(ds)[np.argsort(np.concatenate(seqs), kind="stable")].tolist()
    return order, kept, ds


def parseRawBedpe2(fs, fout, cs, cut, logger):
    """cLoops/io.py:132-189 + txt2jd (:192-203) in one step: per-chromosome ``[id, cA, cB]`` int64
    matrices (id restarts at 0 per chromosome, rows in file order) written straight to ``.jd``.
    Returns the list of .jd paths in order of first appearance."""
    order, per, _ = readBedpe(fs, cs, cut, logger)
    return _write_jd(fout, per, order)


def parseRawBedpe(fs, fout, cs, cut, logger):
    """cLoops/io.py:62-129: as parseRawBedpe2 but drops duplicate (cA, cB) per chromosome (first one
    wins) and collects the distances of opposite-strand PETs (input of estFragSize when eps is auto)."""
    order, per, ds = readBedpe(fs, cs, cut, logger, dedup=True)
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
