"""Drop-in for ``cLoops.pipe`` (cLoops/pipe.py): per-chromosome dispatch of clustering and scoring,
the round loop with its distance cut-off feedback, and the ``cLoops`` command line.

Where the reference fans chromosomes out to joblib worker processes (pipe.py:117,184), this module
keeps every chromosome resident in HBM for the whole run and drives the CUDA kernels from one host
process per GPU; with several ranks (torchrun) chromosomes are sharded across GPUs (``dist.py``) and
only the per-round distance lists and the final loop table travel between ranks.
"""
from __future__ import annotations

import os
import shutil
import sys
from datetime import datetime

import numpy as np
import pandas as pd

from . import _lib, device, dist
from .cModel import getIntSig, markIntSig, markIntSigHic
from .ests import cut_from_moments, estFragSize, estIntSelCutFrag
from .io import loops2juice, loops2washU, parseJd, parseRawBedpe, parseRawBedpe2
from .utils import getLogger, mainHelp

logger = None

#: which clusterer ``singleDBSCAN`` runs: the reference binds cDBSCAN2 (pipe.py:42) and keeps
#: blockDBSCAN as a commented alternative (pipe.py:43)
DBSCAN_VARIANT = _lib.V2


class _Resident:
    """One chromosome's PETs on the host and in HBM, loaded once per .jd path."""
    _cache: dict = {}

    def __init__(self, f):
        self.key, mat = parseJd(f, cut=0)
        self.ids = np.ascontiguousarray(mat[:, 0]) if len(mat) else np.zeros(0, np.int64)
        self.X = np.ascontiguousarray(mat[:, 1]) if len(mat) else np.zeros(0, np.int64)
        self.Y = np.ascontiguousarray(mat[:, 2]) if len(mat) else np.zeros(0, np.int64)
        self.dx = device.to_device_i32(self.X, "X")
        self.dy = device.to_device_i32(self.Y, "Y")

    @classmethod
    def get(cls, f):
        st = os.stat(f)
        tag = (os.path.abspath(f), st.st_mtime_ns, st.st_size)
        hit = cls._cache.get(tag[0])
        if hit is None or hit[0] != tag:
            hit = (tag, cls(f))
            cls._cache[tag[0]] = hit
        return hit[1]

    @classmethod
    def clear(cls):
        cls._cache.clear()


def _single(f, eps, minPts, cut=0):
    """One chromosome of one round -> (key, f, dataI, dataS, dis, dss) with dis/dss as float64 arrays.
    The cut filter, the clusterer and the per-cluster reduction all run on the GPU; the host receives
    the candidate records and one kind byte per PET."""
    ch = _Resident.get(f)
    key = ch.key
    dataI, dataS = [], []
    dis = np.zeros(0, np.float64)
    dss = []
    d = ch.Y - ch.X
    if cut > 0:
        dss.append(d[d < cut].astype(np.float64))
        n_act = int((d >= cut).sum())
    else:
        n_act = len(d)
    if n_act == 0:
        return key, f, dataI, dataS, dis, (np.concatenate(dss) if dss else np.zeros(0, np.float64))
    sys.stderr.write("Clustering %s and %s using eps as %s, minPts as %s,pre-set distance cutoff as > %s\n" %
                     (key[0], key[1], eps, minPts, cut))
    c = device.cluster_and_summarise(ch.dx, ch.dy, int(eps), int(minPts), DBSCAN_VARIANT, int(cut) if cut > 0 else 0)
    bbox, kind, row_kind = c.bbox.cpu().numpy(), c.kind.cpu().numpy(), c.row_kind.cpu().numpy()
    for b in bbox[kind == 1].tolist():
        dataI.append([key[0], b[0], b[1], key[1], b[2], b[3]])
    for b in bbox[kind == 2].tolist():
        dataS.append([key[0], b[0], b[1], key[1], b[2], b[3]])
    in_i, in_s = row_kind == 1, row_kind == 2
    sys.stderr.write("Clustering %s and %s finished. Estimated %s self-ligation reads and %s inter-ligation reads\n" %
                     (key[0], key[1], int(in_s.sum()), int(in_i.sum())))
    if len(dataI) > 0:
        dis = d[in_i].astype(np.float64)
    if len(dataS) > 0:
        dss.append(d[in_s].astype(np.float64))
    return key, f, dataI, dataS, dis, (np.concatenate(dss) if dss else np.zeros(0, np.float64))


def _moments(vals):
    """(n, mean, M2) of log2(vals) in float64 on the device; vals: positive int32 CUDA tensor."""
    import torch
    n = int(vals.numel())
    if n == 0:
        return (0, 0.0, 0.0)
    x = torch.log2(vals.to(torch.float64))
    mean = x.mean()
    return (n, float(mean), float(((x - mean) ** 2).sum()))


def _single_stats(f, eps, minPts, cut=0):
    """As _single, but the distance collections stay in HBM: returns the candidate records, the raw
    sizes of dis / dss, their log2 moments and the positive self-ligation distances (device tensor)."""
    import torch
    ch = _Resident.get(f)
    key = ch.key
    dataI, dataS = [], []
    dd = ch.dy - ch.dx                                   # Y - X on the device
    empty = torch.zeros(0, dtype=torch.int32, device=dd.device)
    removed = dd[dd < cut] if cut > 0 else empty
    n_act = int(dd.numel() - removed.numel())
    if n_act == 0:
        pos = removed.abs()
        pos = pos[pos > 0]
        return key, f, dataI, dataS, 0, int(removed.numel()), (0, 0.0, 0.0), _moments(pos), pos
    sys.stderr.write("Clustering %s and %s using eps as %s, minPts as %s,pre-set distance cutoff as > %s\n" %
                     (key[0], key[1], eps, minPts, cut))
    p = device.Pass(ch.dx, ch.dy, int(eps), int(minPts), DBSCAN_VARIANT, int(cut) if cut > 0 else 0, score=False)
    bbox_h, kind_h = p.bbox.cpu().numpy(), p.kind.cpu().numpy()
    members, member_kind = p.ys - p.xs, p.member_kind         # Y - X of every clustered PET, its cluster's kind
    for b in bbox_h[kind_h == 1].tolist():
        dataI.append([key[0], b[0], b[1], key[1], b[2], b[3]])
    for b in bbox_h[kind_h == 2].tolist():
        dataS.append([key[0], b[0], b[1], key[1], b[2], b[3]])
    inter = members[member_kind == 1] if dataI else empty
    selfm = members[member_kind == 2] if dataS else empty
    sys.stderr.write("Clustering %s and %s finished. Estimated %s self-ligation reads and %s inter-ligation reads\n" %
                     (key[0], key[1], int(selfm.numel()), int(inter.numel())))
    n_dis, n_dss = int(inter.numel()), int(removed.numel() + selfm.numel())
    inter = inter.abs()
    inter = inter[inter > 0]
    selfd = torch.cat([removed, selfm]).abs()
    selfd = selfd[selfd > 0]
    p.close()
    return key, f, dataI, dataS, n_dis, n_dss, _moments(inter), _moments(selfd), selfd


def _round(fs, eps, minPts, cut):
    """One clustering round over all chromosomes with the cut-off statistics reduced on the GPU.
    -> (dataI, dataS, n_dis, n_dss, cut_or_None)"""
    import torch
    mine = dist.my_share(fs)
    full = {f: _single_stats(f, eps, minPts, cut) for f in mine}
    part = {f: r[:8] for f, r in full.items()}           # small host objects travel through the object gather
    ds = dist.merge_in_order(fs, part)
    dataI, dataS, n_dis, n_dss, mi, ms = {}, [], 0, 0, [], []
    used = set()
    for f, d in zip(fs, ds):
        if len(d[2]) == 0:                               # pipe.py:121-122: chromosomes without inter-ligation
            continue                                     # clusters contribute nothing, not even their dss
        used.add(f)
        dataI[d[0]] = {"f": d[1], "records": d[2]}
        dataS.extend(d[3])
        n_dis += d[4]
        n_dss += d[5]
        mi.append(d[6])
        ms.append(d[7])
    if len(dataI) == 0 or n_dis == 0 or n_dss == 0:
        return dataI, dataS, n_dis, n_dss, None
    local = [full[f][8] for f in mine if f in used]
    if local:
        local = torch.cat(local)
    else:                                                 # this rank owns no contributing chromosome
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        local = torch.zeros(0, dtype=torch.int32, device=dev)
    allself = dist.all_gather_concat(local)
    srt = torch.sort(allself).values
    k = int(srt.numel())
    mid = srt[[(k - 1) // 2, k // 2]].cpu().tolist()
    cut_2, frags = cut_from_moments(mi, ms, _TwoMiddle(k, mid))
    return dataI, dataS, n_dis, n_dss, cut_2


class _TwoMiddle:
    """What cut_from_moments needs of the sorted self-ligation distances: length and the two middle values."""

    def __init__(self, k, mid):
        self.k, self.mid = k, mid

    def __len__(self):
        return self.k

    def __getitem__(self, i):
        return self.mid[0] if i == (self.k - 1) // 2 else self.mid[1]


def singleDBSCAN(f, eps, minPts, cut=0):
    """cLoops/pipe.py:52-110 for one chromosome -> ``(key, f, dataI, dataS, dis, dss)`` (lists, as the
    reference returns them)."""
    key, f, dataI, dataS, dis, dss = _single(f, eps, minPts, cut)
    return key, f, dataI, dataS, dis.tolist(), dss.tolist()


def runDBSCAN(fs, eps, minPts, cut=0, cpu=1):
    """cLoops/pipe.py:113-127.  Chromosomes owned by this rank are clustered here; results of all ranks
    are merged in file order so every rank returns what the reference's parent process would.  The
    distance collections come back as float64 arrays (the reference returns lists; its only consumer,
    pipe(), wraps them in np.array, pipe.py:259)."""
    mine = dist.my_share(fs)
    part = {f: _single(f, eps, minPts, cut) for f in mine}
    ds = dist.merge_in_order(fs, part)
    dataI, dataS, dis, dss = {}, [], [], []
    for d in ds:
        if len(d[2]) == 0:
            continue
        dataI[d[0]] = {"f": d[1], "records": d[2]}
        dataS.extend(d[3])
        dis.append(d[4])
        dss.append(d[5])
    dis = np.concatenate(dis) if dis else np.zeros(0, np.float64)
    dss = np.concatenate(dss) if dss else np.zeros(0, np.float64)
    return dataI, dataS, dis, dss


def filterClusterByDis(data, cut):
    """cLoops/pipe.py:130-143: keep inter-ligation clusters whose anchor-centre distance is >= cut."""
    for key in data:
        data[key]["records"] = [r for r in data[key]["records"] if (r[4] + r[5]) // 2 - (r[1] + r[2]) // 2 >= cut]
    return data


def checkSameLoop(ra, rb):
    """cLoops/pipe.py:146-152."""
    return ra[1] == rb[1] and ra[2] == rb[2] and ra[4] == rb[4] and ra[5] == rb[5]


def combineTwice(dataI, dataI_2):
    """cLoops/pipe.py:155-174: append the records of a new round unless the exact bbox is already known."""
    for key in dataI_2.keys():
        if key not in dataI:
            dataI[key] = {"f": dataI_2[key]["f"], "records": dataI_2[key]["records"]}
            continue
        known = set((r[1], r[2], r[4], r[5]) for r in dataI[key]["records"])
        for r in dataI_2[key]["records"]:
            if (r[1], r[2], r[4], r[5]) not in known:
                dataI[key]["records"].append(r)
    return dataI


def runStat(dataI, minPts, cut, cpu, fout, hichip=0):
    """cLoops/pipe.py:177-203 -> 0 on success, 1 when no loop survives."""
    _log().info("Starting estimate significance for interactions using distance cutoff as %s" % cut)
    keys = list(dataI.keys())
    mine = dist.my_share(keys, weights=[len(dataI[k]["records"]) for k in keys])
    part = {k: getIntSig(dataI[k]["f"], dataI[k]["records"], minPts, cut) for k in mine}
    ds = [d for d in dist.merge_in_order(keys, part) if d is not None]
    if len(ds) == 0:
        _log().error("Something wrong, no loops found, sorry, bye.")
        return 1
    ds = pd.concat(ds)
    if dist.rank() != 0:
        return 0
    try:
        ds = markIntSigHic(ds) if hichip else markIntSig(ds)
        ds.to_csv(fout + ".loop", sep="\t", index_label="loopId")
    except Exception:
        _log().warning("Something wrong happend to significance estimation, only output called loops")
        ds.to_csv(fout + "_raw.loop", sep="\t", index_label="loopId")
    return 0


def _log():
    global logger
    if logger is None:
        logger = getLogger(os.path.join(os.getcwd(), "cLoops.log"))
    return logger


def pipe(fs, fout, eps, minPts, chroms="", cpu=1, tmp=0, hic=0, washU=0, juice=0, cut=0, plot=0, max_cut=False):
    """cLoops/pipe.py:206-295."""
    log = _log()
    chroms = [] if chroms == "" else set(chroms.split(","))
    # rank 0 alone looks at / creates the output directory and every rank follows its decision: a rank that
    # checked isdir after rank 0's mkdir would otherwise leave while the others wait in a collective
    ok = True
    if dist.rank() == 0:
        if os.path.isdir(fout):
            log.error("working directory %s exists, return." % fout)
            ok = False
        else:
            os.mkdir(fout)
    if not dist.broadcast_object(ok):
        return
    if dist.rank() == 0:
        if eps == 0:
            cfs, ds = parseRawBedpe(fs, fout, chroms, cut, log)
        else:
            cfs, ds = parseRawBedpe2(fs, fout, chroms, cut, log), None
    else:
        cfs, ds = None, None
    cfs, ds = dist.broadcast_object((cfs, ds))
    if eps == 0:
        eps = [estFragSize(ds) * 2]
    dataI = {}
    cuts = [cut]
    for ep in eps:
        for m in minPts:
            dataI_2, dataS_2, n_dis, n_dss, cut_2 = _round(cfs, ep, m, cut)
            if len(dataI_2) == 0:
                log.info("ERROR: no inter-ligation PETs detected for eps %s minPts %s,can't model the distance cutoff,continue anyway" % (ep, m))
                continue
            if cut_2 is None:
                dataI = combineTwice(dataI, dataI_2)
                continue
            log.info("Estimated inter-ligation and self-ligation distance cutoff as %s for eps=%s,minPts=%s" % (cut_2, ep, m))
            cuts.append(cut_2)
            cut = cut_2
            dataI = combineTwice(dataI, dataI_2)
    cuts = [c for c in cuts if c > 0]
    cut = np.max(cuts) if max_cut else np.min(cuts)
    dataI = filterClusterByDis(dataI, cut)
    e = runStat(dataI, minPts, 0, cpu, fout, hic)
    _Resident.clear()
    dist.barrier()
    if dist.rank() != 0:
        return
    if e:
        shutil.rmtree(fout)
        return
    if washU:
        loops2washU(fout + ".loop", fout + "_loops_washU.txt", log)
    if juice:
        loops2juice(fout + ".loop", fout + "_loops_juicebox.txt", log)
    if not tmp:
        shutil.rmtree(fout)


def _int_list(v, reverse):
    """pipe.py:308-328 parses ``-eps`` / ``-minPts``: comma list -> sorted ints; single value -> [int] (0 stays 0)."""
    if "," in str(v):
        return sorted((int(x) for x in str(v).split(",")), reverse=reverse)
    v = int(v)
    return [v] if v != 0 else 0


def main(argv=None):
    """cLoops/pipe.py:298-352 (console script ``cLoops``)."""
    global logger
    start = datetime.now()
    logger = getLogger(os.path.join(os.getcwd(), "cLoops.log"))
    op = mainHelp(argv)
    logger.info("Command line: cLoops -f {} -o {} -m {} -eps {} -minPts {} -p {} -w {} -j {} -s {} -c {} -hic {} -cut {} -plot {} -max_cut {}".format(
        op.fnIn, op.fnOut, op.mode, op.eps, op.minPts, op.cpu, op.washU, op.juice, op.tmp, op.chroms, op.hic, op.cut,
        op.plot, op.max_cut))
    presets = {1: ([500, 1000, 2000], [5], 0), 2: ([1000, 2000, 5000], [5], 0),
               3: ([5000, 7500, 10000], [50, 40, 30, 20], 1), 4: ([2500, 5000, 7500, 10000], [30, 20], 1)}
    if op.mode == 0:
        eps = _int_list(op.eps, reverse=False)
        minPts = _int_list(op.minPts, reverse=True)
        if minPts == 0:
            logger.error("minPts not assigned!")
            return
        hic = op.hic
    else:
        eps, minPts, hic = presets[op.mode]
    logger.info("mode:%s\t eps:%s\t minPts:%s\t hic:%s\t" % (op.mode, eps, minPts, hic))
    dist.init_from_env()
    pipe(op.fnIn.split(","), op.fnOut, eps, minPts, op.chroms, op.cpu, op.tmp, hic, op.washU, op.juice, op.cut,
         op.plot, op.max_cut)
    dist.shutdown()
    logger.info("cLoops finished. Used CPU time: %s Bye!\n\n\n" % (datetime.now() - start))


if __name__ == "__main__":
    main()
