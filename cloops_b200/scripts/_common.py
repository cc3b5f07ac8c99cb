"""Pieces shared by quantifyLoops and deLoops (both scripts carry their own copy in the reference)."""
from __future__ import annotations

import os

import numpy as np

from ..io import parseIv


def loop_columns(header: str, ivac=None, ivbc=None):
    """Positions of the iva / ivb columns.  The reference hard-codes 6 and 7 (scripts/deLoops:34,
    scripts/quantifyLoops.py:92), the alphabetical column order old pandas gave the .loop table; a table
    written by a current pandas keeps insertion order, so the header decides when it names them."""
    cols = header.rstrip("\n").split("\t")
    if ivac is None:
        ivac = cols.index("iva") if "iva" in cols else 6
    if ivbc is None:
        ivbc = cols.index("ivb") if "ivb" in cols else 7
    return ivac, ivbc


def preDs(f, d, chroms=(), ivac=None, ivbc=None, logger=None):
    """scripts/deLoops:34-70 = scripts/quantifyLoops.py:92-128: significant loops (last column >= 1) of a
    .loop file, grouped by chromosome, with the .jd file of each chromosome in directory d."""
    records = {}
    for c in chroms:
        records[c] = {"rs": {}, "f": ""}
    with open(f) as fh:
        for i, line in enumerate(fh):
            if i == 0:
                ivac, ivbc = loop_columns(line, ivac, ivbc)
                continue
            line = line.split("\n")[0].split("\t")
            if float(line[-1]) < 1:
                continue
            iva, ivb = parseIv(line[ivac]), parseIv(line[ivbc])
            if len(chroms) > 0 and iva[0] not in chroms:
                continue
            if iva[0] not in records:
                records[iva[0]] = {"rs": {}, "f": ""}
            records[iva[0]]["rs"][line[0]] = iva + ivb
    for chrom in list(records.keys()):
        if len(records[chrom]["rs"]) == 0:
            del records[chrom]
            continue
        jd = os.path.join(d, "%s-%s.jd" % (chrom, chrom))
        if os.path.isfile(jd):
            records[chrom]["f"] = jd
        else:
            if logger is not None:
                logger.warning("%s not found, however there are loops in that chromosome." % jd)
            del records[chrom]
    return records


def loop_intervals(rs):
    """keys, chromosomes and the int64 [m,4] array (iva0, iva1, ivb0, ivb1) of a chromosome's loops, file order."""
    keys = list(rs.keys())
    iv = np.array([[rs[k][1], rs[k][2], rs[k][4], rs[k][5]] for k in keys], dtype=np.int64).reshape(-1, 4)
    return keys, [rs[k][0] for k in keys], iv


def nearby_pairs(iv, win=5):
    """The 2*win x 2*win shifted window pairs of every loop (cModel.py:83-105, py2 integer arithmetic):
    int64 [m, (2*win)^2, 4], pair (i, j) = (A_i, B_j) at position i * 2*win + j."""
    iv = np.asarray(iv, dtype=np.int64).reshape(-1, 4)
    ca, cb = (iv[:, 0] + iv[:, 1]) // 2, (iv[:, 2] + iv[:, 3]) // 2
    sa, sb = (iv[:, 1] - iv[:, 0]) // 2, (iv[:, 3] - iv[:, 2]) // 2
    step = (sa + sb) // 2
    shifts = np.array([i for i in range(-win, win + 1) if i != 0], dtype=np.int64)
    a0 = np.maximum(0, ca[:, None] + shifts[None, :] * step[:, None] - sa[:, None])
    a1 = np.maximum(0, ca[:, None] + shifts[None, :] * step[:, None] + sa[:, None])
    b0 = np.maximum(0, cb[:, None] + shifts[None, :] * step[:, None] - sb[:, None])
    b1 = np.maximum(0, cb[:, None] + shifts[None, :] * step[:, None] + sb[:, None])
    k = len(shifts)
    out = np.empty((iv.shape[0], k, k, 4), dtype=np.int64)
    out[:, :, :, 0] = a0[:, :, None]
    out[:, :, :, 1] = a1[:, :, None]
    out[:, :, :, 2] = b0[:, None, :]
    out[:, :, :, 3] = b1[:, None, :]
    return out.reshape(iv.shape[0], k * k, 4)
