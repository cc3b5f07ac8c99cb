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
from . import cModel
from .cModel import getIntSig, markIntSig, markIntSigHic
from .ests import cut_from_round, estFragSize, estIntSelCutFrag
from .io import loops2juice, loops2washU, parseJd, parseRawBedpe, parseRawBedpe2, readBedpe
from .utils import getLogger, mainHelp

logger = None


def _resident_model(f):
    """cModel.getGenomeCoverage hook: the chromosome already resident in HBM (discut = 0), instead of a second parse + upload."""
    hit = _Resident._cache.get(f if f.startswith("mem:") else os.path.abspath(f))
    return hit[1] if hit is not None else None


#: which clusterer ``singleDBSCAN`` runs: the reference binds cDBSCAN2 (pipe.py:42) and keeps
#: blockDBSCAN as a commented alternative (pipe.py:43)
DBSCAN_VARIANT = _lib.V2


class _Resident:
    """One chromosome's PETs on the host and in HBM, loaded once per .jd path -- or registered from arrays under a
    pseudo path (``_Resident.register``), which is how bench.py and the tests feed synthetic chromosomes without files."""
    _cache: dict = {}

    def __init__(self, f=None, key=None, X=None, Y=None, ids=None, dx=None, dy=None):
        if f is not None:
            self.key, mat = parseJd(f, cut=0)
            ids = np.ascontiguousarray(mat[:, 0]) if len(mat) else np.zeros(0, np.int64)
            X = np.ascontiguousarray(mat[:, 1]) if len(mat) else np.zeros(0, np.int64)
            Y = np.ascontiguousarray(mat[:, 2]) if len(mat) else np.zeros(0, np.int64)
        else:
            self.key = tuple(key)
        self.ids, self.X, self.Y = ids, X, Y
        self.n = len(X)
        self.dx = device.to_device_i32(X, "X") if dx is None else dx
        self.dy = device.to_device_i32(Y, "Y") if dy is None else dy
        self._base = None                  # (eps, device.Index built with cut = 0): shared by the rounds of one eps

    def base_index(self, eps):
        """The chromosome's full (cut = 0) index for ``eps``, built on first use and kept until another eps is asked for."""
        if self._base is not None and self._base[0] != eps:
            self._base[1].close()
            self._base = None
        if self._base is None:
            self._base = (eps, device.Index(self.dx, self.dy, int(eps), 0))
        return self._base[1]

    def drop_base(self):
        if self._base is not None:
            self._base[1].close()
            self._base = None

    @classmethod
    def get(cls, f):
        hit = cls._cache.get(f if f.startswith("mem:") else os.path.abspath(f))
        if hit is not None and hit[0] is None:
            return hit[1]
        st = os.stat(f)
        tag = (os.path.abspath(f), st.st_mtime_ns, st.st_size)
        if hit is None or hit[0] != tag:
            hit = (tag, cls(f))
            cls._cache.pop(tag[0], None)
            cls._cache[tag[0]] = hit
            cls._evict(keep=tag[0])
        return hit[1]

    #: file-backed chromosomes kept resident: at most this many bytes of coordinates in HBM (least recently loaded go first),
    #: so direct callers of singleDBSCAN / runDBSCAN do not accumulate device memory; pipe() clears the cache when it is done
    BUDGET_BYTES = int(os.environ.get("CLOOPS_RESIDENT_BYTES", str(48 << 30)))

    @classmethod
    def _evict(cls, keep):
        total = sum(v[1].n * 8 for v in cls._cache.values())
        for name in list(cls._cache):
            if total <= cls.BUDGET_BYTES:
                break
            if name == keep or cls._cache[name][0] is None:          # registered arrays belong to their owner
                continue
            total -= cls._cache.pop(name)[1].n * 8

    @classmethod
    def register(cls, name, X, Y, dx=None, dy=None):
        """A chromosome given as host arrays (and, optionally, its int32 CUDA tensors already in HBM) -> pseudo path
        ``mem:<name>-<name>.jd``."""
        f = "mem:%s-%s.jd" % (name, name)
        cls._cache[f] = (None, cls(None, (name, name), X, Y, None, dx, dy))
        return f

    @classmethod
    def clear(cls):
        for _, ch in cls._cache.values():
            ch.drop_base()
        cls._cache.clear()


cModel.RESIDENT = _resident_model


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


class _RoundAcc:
    """Device accumulators of one clustering round: exact histogram of the positive self-ligation distances and the
    log2 moments of both distance collections (``cloops_pass_run_stats``, csrc/roundstats.cu).  They replace the pooled
    ``dis`` / ``dss`` lists of cLoops/pipe.py:120-127; across GPUs they are summed with one all-reduce per round."""
    _per_device: dict = {}

    def __init__(self, dev):
        import torch
        self.hist = torch.zeros(_lib.ROUND_HIST_BINS + 1, dtype=torch.int32, device=dev)
        self.mom = torch.zeros(_lib.ROUND_MOM, dtype=torch.float64, device=dev)

    @classmethod
    def get(cls):
        dev = device.require_cuda()
        if dev not in cls._per_device:
            cls._per_device[dev] = cls(dev)
        return cls._per_device[dev]

    def reset(self):
        self.hist.zero_()
        self.mom.zero_()

    def reduce(self):
        dist.all_reduce_sum(self.hist)
        dist.all_reduce_sum(self.mom)

    def middle(self):
        """-> (lo, hi, mom): the two middle order statistics of the self-ligation distances and the moments (host)."""
        import ctypes as C
        import torch
        mid, mom = (C.c_int64 * 2)(), (C.c_double * _lib.ROUND_MOM)()
        _lib.check(_lib.lib().cloops_round_middle(self.hist.data_ptr(), self.mom.data_ptr(), C.addressof(mid), C.addressof(mom),
                                                  torch.cuda.current_stream().cuda_stream))
        return int(mid[0]), int(mid[1]), [float(v) for v in mom]


#: set per run by _rounds: an eps clustered with several minPts values (presets -m 3 / -m 4) builds the chromosome's index once
#: (cut = 0) and derives every round's index from it by a compaction instead of a sort (cloops_pass_run_base)
REUSE_INDEX = False


def _weights(fs):
    """PETs per chromosome file (LPT packing of chromosomes onto ranks, dist.assign): a .jd holds 24 B per PET."""
    out = []
    for f in fs:
        hit = _Resident._cache.get(f)
        out.append(hit[1].n * 24 if hit is not None else (os.path.getsize(f) if os.path.exists(f) else 1))
    return out


def _cluster_chrom(f, eps, minPts, cut, acc):
    """One chromosome of one round on the GPU: cut filter (pipe.py:59-63), clusterer (:70), per-cluster records (:76-102)
    and this chromosome's share of the round's distance statistics (:106-109), in ONE C-ABI call.
    -> (key, inter-ligation records int32 [K,4] = minX, maxX, minY, maxY in ascending cluster id, #self-ligation clusters)"""
    ch = _Resident.get(f)
    key = ch.key
    empty = np.zeros((0, 4), np.int32)
    if ch.n == 0:
        return key, empty, 0
    sys.stderr.write("Clustering %s and %s using eps as %s, minPts as %s,pre-set distance cutoff as > %s\n" %
                     (key[0], key[1], eps, minPts, cut))
    base = ch.base_index(int(eps)) if (REUSE_INDEX and DBSCAN_VARIANT != _lib.BLOCK) else None
    p = device.Pass(ch.dx, ch.dy, int(eps), int(minPts), DBSCAN_VARIANT, int(cut) if cut > 0 else 0, score=False,
                    stats=(acc.hist, acc.mom), base=base)
    bbox, size, kind = p.records()
    p.close()
    sys.stderr.write("Clustering %s and %s finished. Estimated %s self-ligation reads and %s inter-ligation reads\n" %
                     (key[0], key[1], int(size[kind == 2].sum()), int(size[kind == 1].sum())))
    return key, bbox[kind == 1], int((kind == 2).sum())


#: chromosomes of a round in flight at once on this GPU, each on its own CUDA stream and host thread: while one pass waits
#: for a size it needs on the host (five short synchronisations per pass), the kernels of another keep the SMs busy
STREAMS = max(1, int(os.environ.get("CLOOPS_STREAMS", "6")))
LARGEST_FIRST = os.environ.get("CLOOPS_LARGEST_FIRST", "1") != "0"
_pool = {}


def _on_streams(items, fn, sizes=None):
    """[fn(item) for item in items], up to STREAMS of them in flight, each host thread on a CUDA stream of its own (the
    streams wait for the caller's stream first and the caller's stream waits for them afterwards).  With ``sizes`` the
    largest items start first, so that the call does not end on one large item running alone (CLOOPS_LARGEST_FIRST=0:
    list order); the results come back in list order either way."""
    import torch
    items = list(items)
    if STREAMS <= 1 or len(items) <= 1 or not torch.cuda.is_available():
        return [fn(it) for it in items]
    if sizes is not None and LARGEST_FIRST:
        order = sorted(range(len(items)), key=lambda k: -sizes[k])
        back = _on_streams([items[k] for k in order], fn)
        out = [None] * len(items)
        for k, r in zip(order, back):
            out[k] = r
        return out
    from concurrent.futures import ThreadPoolExecutor
    dev = torch.cuda.current_device()
    if "ex" not in _pool:
        _pool["ex"] = ThreadPoolExecutor(max_workers=STREAMS)
        _pool["streams"] = {}
    main = torch.cuda.current_stream()

    def work(it):
        import threading
        torch.cuda.set_device(dev)
        tid = threading.get_ident()
        st = _pool["streams"].get((dev, tid))
        if st is None:
            st = _pool["streams"][(dev, tid)] = torch.cuda.Stream(device=dev)
        st.wait_stream(main)
        with torch.cuda.stream(st):
            out = fn(it)
        return out, st

    res = list(_pool["ex"].map(work, items))
    for _, st in res:
        main.wait_stream(st)
    return [r for r, _ in res]


def _cluster_many(files, eps, minPts, cut, acc):
    """_cluster_chrom for every file, results in file order."""
    if not acc.hist.is_cuda:
        return [_cluster_chrom(f, eps, minPts, cut, acc) for f in files]
    return _on_streams(files, lambda f: _cluster_chrom(f, eps, minPts, cut, acc), sizes=[_Resident.get(f).n for f in files])


def _round(fs, eps, minPts, cut, weights=None):
    """One clustering round over the chromosomes this rank owns, the distance cut-off reduced on the GPUs.
    -> (dataI_2, n_self_clusters, len(dis), len(dss), cut_2 or None, n_chromosomes_with_inter_ligation_clusters);
    dataI_2 = {key: {"f": f, "records": int32 array [K,4]}} holds THIS rank's chromosomes only; the four numbers after it
    are global (all ranks)."""
    acc = _RoundAcc.get()
    acc.reset()
    dataI, n_self = {}, 0
    mine = dist.my_share(fs, _weights(fs) if weights is None else weights)
    for f, (key, inter, ns) in zip(mine, _cluster_many(mine, eps, minPts, cut, acc)):
        if len(inter) == 0:                                # pipe.py:121-122
            continue
        dataI[key] = {"f": f, "records": inter}
        n_self += ns
    acc.reduce()
    lo, hi, mom = acc.middle()
    n_dis, n_dss, n_contrib = int(mom[6]), int(mom[7]), int(mom[8])
    if n_contrib == 0 or n_dis == 0 or n_dss == 0:
        return dataI, n_self, n_dis, n_dss, None, n_contrib
    cut_2 = cut_from_round(mom, lo, hi)
    if cut_2 is None:
        # the median left the histogram range, or 2**cut sits on an integer boundary where the summation order of the
        # moments could flip int(): pool the distances on the host and run the reference's numpy estimate instead
        _, _, dis, dss = runDBSCAN(fs, eps, minPts, cut, _weights=weights)
        cut_2 = estIntSelCutFrag(dis, dss)[0]
    return dataI, n_self, n_dis, n_dss, cut_2, n_contrib


def _combine_rounds(rounds):
    """combineTwice (pipe.py:155-174) over all rounds of one chromosome at once: a record is dropped iff the exact same
    bbox was produced by an EARLIER round; order = round order, then cluster id.  rounds: list of int32 [K,4] arrays.
    The membership test runs in host C++ (``cloops_combine_rounds``: one pass over a hash table)."""
    if len(rounds) == 1:
        return rounds[0]
    allr = np.ascontiguousarray(np.concatenate(rounds), dtype=np.int32)
    rnd = np.ascontiguousarray(np.repeat(np.arange(len(rounds), dtype=np.int32), [len(r) for r in rounds]))
    keep = np.empty(len(allr), np.uint8)
    _lib.check(_lib.lib().cloops_combine_rounds(allr.ctypes.data, rnd.ctypes.data, len(allr), keep.ctypes.data))
    return allr[keep.view(bool)]


def _finalize_records(e, cut):
    """combineTwice over the chromosome's rounds + filterClusterByDis (pipe.py:155-174,130-143) -> e["records"] int64 [K,4]."""
    if "rounds" in e:
        r = _combine_rounds(e.pop("rounds")).astype(np.int64)
        e["records"] = r[(r[:, 2] + r[:, 3]) // 2 - (r[:, 0] + r[:, 1]) // 2 >= cut]
    return e


def _rounds(cfs, eps, minPts, cut, max_cut, log, weights=None, finalize=True):
    """The round loop of cLoops/pipe.py:247-281: clustering rounds with the distance cut-off fed forward, candidates of all
    rounds merged (combineTwice) and filtered by the final cut-off.  -> (dataI of this rank's chromosomes with records as
    int64 arrays [K,4] and "first" = first round that produced the chromosome, final cut)"""
    global REUSE_INDEX
    REUSE_INDEX = len(minPts) > 1 and os.environ.get("CLOOPS_REUSE_INDEX", "1") != "0"
    dataI, cuts, rnd = {}, [cut], 0
    for ep in eps:
        for m in minPts:
            d2, n_self, n_dis, n_dss, cut_2, n_contrib = _round(cfs, ep, m, cut, weights)
            if n_contrib == 0:
                log.info("ERROR: no inter-ligation PETs detected for eps %s minPts %s,can't model the distance cutoff,continue anyway" % (ep, m))
                continue
            for key, v in d2.items():
                dataI.setdefault(key, {"f": v["f"], "rounds": [], "first": rnd})["rounds"].append(v["records"])
            rnd += 1
            if cut_2 is None:
                continue
            log.info("Estimated inter-ligation and self-ligation distance cutoff as %s for eps=%s,minPts=%s" % (cut_2, ep, m))
            cuts.append(cut_2)
            cut = cut_2
    for hit in _Resident._cache.values():                      # the per-eps base indexes are not needed for scoring
        hit[1].drop_base()
    cuts = [c for c in cuts if c > 0]
    cut = np.max(cuts) if max_cut else np.min(cuts)
    if finalize:
        for e in dataI.values():
            _finalize_records(e, cut)
    return dataI, cut


def singleDBSCAN(f, eps, minPts, cut=0):
    """cLoops/pipe.py:52-110 for one chromosome -> ``(key, f, dataI, dataS, dis, dss)`` (lists, as the
    reference returns them)."""
    key, f, dataI, dataS, dis, dss = _single(f, eps, minPts, cut)
    return key, f, dataI, dataS, dis.tolist(), dss.tolist()


def runDBSCAN(fs, eps, minPts, cut=0, cpu=1, _weights=None):
    """cLoops/pipe.py:113-127.  Chromosomes owned by this rank are clustered here; results of all ranks
    are merged in file order so every rank returns what the reference's parent process would.  The
    distance collections come back as float64 arrays (the reference returns lists; its only consumer,
    pipe(), wraps them in np.array, pipe.py:259)."""
    mine = dist.my_share(fs, _weights)
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


def runStat(dataI, minPts, cut, cpu, fout, hichip=0, _local=False):
    """cLoops/pipe.py:177-203 -> 0 on success, 1 when no loop survives.  ``_local`` (pipe()): ``dataI`` holds only the
    chromosomes resident on this rank; every rank scores its own and rank 0 concatenates the tables in the reference's
    order (dict insertion order: first round that produced the chromosome, then file order)."""
    ds = _score(dataI, minPts, cut, _local)
    if ds is None:
        _log().error("Something wrong, no loops found, sorry, bye.")
        return 1
    if dist.rank() != 0:
        return 0
    try:
        ds = markIntSigHic(ds) if hichip else markIntSig(ds)
        ds.to_csv(fout + ".loop", sep="\t", index_label="loopId")
    except Exception:
        _log().warning("Something wrong happend to significance estimation, only output called loops")
        ds.to_csv(fout + "_raw.loop", sep="\t", index_label="loopId")
    return 0


def _count(dataI, minPts, cut, _local=False):
    """GPU half of runStat: per chromosome the counted candidates (cModel.countCandidates) -> {key: counted} of the
    chromosomes this rank scores."""
    keys = list(dataI.keys())
    mine = keys if _local else dist.my_share(keys, weights=[len(dataI[k]["records"]) for k in keys])
    return {k: cModel.countCandidates(dataI[k]["f"], dataI[k]["records"], minPts, cut) for k in mine}


def _tables(dataI, counted, _local=False, done=False):
    """Host half of runStat: statistics tail per chromosome (``done``: ``counted`` already holds the tables), tables of all
    ranks concatenated in the reference's order (pipe.py:187-191) -> DataFrame on every rank, or None."""
    keys = list(dataI.keys())
    finish = (lambda c: c) if done else cModel.tableFromCounts
    if _local and dist.world() > 1:
        part = {k: (dataI[k].get("first", 0), dataI[k].get("order", 0), finish(counted[k])) for k in keys}
        merged = {}
        for g in dist.all_gather_objects(part):
            merged.update(g)
        ds = [v[2] for _, v in sorted(merged.items(), key=lambda kv: (kv[1][0], kv[1][1])) if v[2] is not None]
    else:
        part = {k: finish(c) for k, c in counted.items()}
        ds = [d for d in dist.merge_in_order(keys, part) if d is not None]
    if len(ds) == 0:
        return None
    return pd.concat(ds)


def _score(dataI, minPts, cut, _local=False):
    """getIntSig over chromosomes (pipe.py:184-191) -> concatenated table on every rank, or None."""
    _log().info("Starting estimate significance for interactions using distance cutoff as %s" % cut)
    keys = list(dataI.keys())
    mine = keys if _local else dist.my_share(keys, weights=[len(dataI[k]["records"]) for k in keys])
    tables = {k: getIntSig(dataI[k]["f"], dataI[k]["records"], minPts, cut) for k in mine}
    return _tables(dataI, tables, _local, done=True)


def call_loops(cfs, eps, minPts, hic=0, cut=0, max_cut=False, weights=None, tail=True, mark=True):
    """pipe() between ingest and output (cLoops/pipe.py:240-284 + 187-196) on chromosomes that are resident in HBM
    (``_Resident.register`` / ``.jd`` paths): the clustering rounds with cut-off feedback, candidate merging and filtering,
    range counts, and -- ``tail=True`` -- the statistics tail and the marked loop table.  ``cfs`` lists ALL chromosomes
    (every rank), ``weights`` their PET counts; each rank works on its own share.
    -> dict(cut, dataI, counted, table)"""
    from concurrent.futures import ThreadPoolExecutor
    import time
    t0 = time.perf_counter()
    dataI, cut = _rounds(cfs, eps, minPts, cut, max_cut, _log(), weights, finalize=False)
    t1 = time.perf_counter()
    for k, f in enumerate(cfs):
        key = tuple(os.path.split(f)[1].replace("mem:", "").replace(".jd", "").split("-"))
        if key in dataI:
            dataI[key]["order"] = k
    out = {"cut": int(cut), "dataI": dataI, "counted": None, "table": None}
    # candidate merging of chromosome k+1 (host C++, releases the GIL) runs while the GPU counts chromosome k; with
    # ``tail`` the statistics tail of a chromosome (host threads) runs while the GPU counts the next ones
    with ThreadPoolExecutor(max_workers=1) as prep, ThreadPoolExecutor(max_workers=max(1, min(6, (os.cpu_count() or 2) // 2))) as ex:
        work = {k: sum(len(r) for r in dataI[k]["rounds"]) if "rounds" in dataI[k] else len(dataI[k]["records"]) for k in dataI}
        first = sorted(dataI, key=lambda k: -work[k]) if LARGEST_FIRST else list(dataI)      # the order the GPU takes them in
        ready = {k: prep.submit(_finalize_records, dataI[k], cut) for k in first}
        futs = {}

        def count_one(k):
            ready[k].result()
            c = cModel.countCandidates(dataI[k]["f"], dataI[k]["records"], minPts, 0)
            if tail:
                futs[k] = ex.submit(cModel.tableFromCounts, c)
            return c

        keys = list(dataI)
        counted = dict(zip(keys, _on_streams(keys, count_one, sizes=[work[k] for k in keys])))
        t2 = time.perf_counter()
        if not tail:
            out["counted"] = counted
            out["host_ms"] = {"rounds": (t1 - t0) * 1e3, "range_counts": (t2 - t1) * 1e3}
            return out
        tables = {k: futs[k].result() for k in keys}
    t3 = time.perf_counter()
    ds = _tables(dataI, tables, _local=True, done=True)
    t4 = time.perf_counter()
    if ds is not None:
        out["table"] = (markIntSigHic(ds) if hic else markIntSig(ds)) if mark else ds
    # wall-clock of the host thread: clustering rounds; range counts (with the statistics tails of the chromosomes counted
    # first running beside them); what is left of the tails once the GPU is done; table gather; significance marks
    out["host_ms"] = {"rounds": (t1 - t0) * 1e3, "range_counts": (t2 - t1) * 1e3, "tail_after_gpu": (t3 - t2) * 1e3,
                      "tables": (t4 - t3) * 1e3, "marks": (time.perf_counter() - t4) * 1e3}
    return out


def finish_loops(run, hic=0):
    """The host tail of call_loops: statistics, de-duplication, tables gathered from all ranks, significance marks."""
    ds = _tables(run["dataI"], run["counted"], _local=True)
    if ds is None:
        return None
    return markIntSigHic(ds) if hic else markIntSig(ds)


def _log():
    global logger
    if logger is None:
        logger = getLogger(os.path.join(os.getcwd(), "cLoops.log"))
    return logger


#: one process that does not keep the temporary files (no ``-s``) skips the .jd round trip; CLOOPS_MEMORY_INGEST=0 writes them
MEMORY_INGEST = os.environ.get("CLOOPS_MEMORY_INGEST", "1") != "0"


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
    if dist.world() == 1 and not tmp and MEMORY_INGEST:
        # the per-chromosome .jd files (pipe.py:231-236) only carry the PETs to the workers and are deleted at the end
        # unless -s is given (pipe.py:293-294): one process that does not keep them goes from the text to HBM directly
        order, per, ds = readBedpe(fs, chroms, cut, log, dedup=(eps == 0))
        cfs = [_Resident.register(c, *per[c]) for c in order]
    else:
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
    log.info("Starting estimate significance for interactions using distance cutoff as 0")
    ds = call_loops(cfs, eps, minPts, hic, cut, max_cut, mark=False)["table"]       # pipe.py:247-284 (+ 184-191)
    e = 0
    if ds is None:
        log.error("Something wrong, no loops found, sorry, bye.")
        e = 1
    elif dist.rank() == 0:                                    # pipe.py:191-202
        try:
            ds = markIntSigHic(ds) if hic else markIntSig(ds)
            ds.to_csv(fout + ".loop", sep="\t", index_label="loopId")
        except Exception:
            log.warning("Something wrong happend to significance estimation, only output called loops")
            ds.to_csv(fout + "_raw.loop", sep="\t", index_label="loopId")
    _Resident.clear()
    _lib.check(_lib.lib().cloops_workspace_release())          # every pass has been fetched: give the scratch blocks back
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
