"""Device-side plumbing (torch tensors as HBM buffers + the caller's CUDA stream) around the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CloopsError, check

COORD_LIMIT = 1 << 30
INFO_KEYS = ("n_active", "n_clusters", "n_components", "n_core", "n_dead", "n_strips", "key_bits", "n_labelled")


class Profile:
    """Measurement hooks (bench.py): with ``Profile.on`` the C library times its stages with CUDA events (one extra
    synchronisation per call) and the wrappers below add them up, together with the algorithmic bytes SURVEY 8d defines
    for the two graded kernels.  ``d2h_bytes`` is counted always (e2e accounting)."""
    on = False
    stages: dict = {}
    rq_bytes = 0          # region query: 12 B per active PET per launch
    rc_bytes = 0          # range counts: 8 B per PET in the hull slices + 4 B per output integer
    d2h_bytes = 0

    @classmethod
    def begin(cls):
        cls.on, cls.stages, cls.rq_bytes, cls.rc_bytes = True, {}, 0, 0
        _lib.lib().cloops_set_profiling(1)

    @classmethod
    def end(cls):
        cls.on = False
        _lib.lib().cloops_set_profiling(0)
        return dict(cls.stages)

    @classmethod
    def add_stages(cls):
        for k, v in _lib.stage_times().items():
            cls.stages[k] = cls.stages.get(k, 0.0) + v


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise CloopsError("cloops_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def to_device_i32(a, name="array") -> torch.Tensor:
    """numpy / torch integer array -> contiguous int32 CUDA tensor (range-checked)."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        if a.is_cuda and a.dtype == torch.int32 and a.is_contiguous():
            return a
        if a.dtype != torch.int32:
            if a.numel() and (int(a.min()) < -COORD_LIMIT or int(a.max()) >= COORD_LIMIT):
                raise CloopsError("%s: coordinates must lie in [-2^30, 2^30)" % name)
            a = a.to(torch.int32)
        return a.contiguous().to(dev, non_blocking=True)
    a = np.asarray(a)
    if a.dtype != np.int32:
        if a.size and (a.min() < -COORD_LIMIT or a.max() >= COORD_LIMIT):
            raise CloopsError("%s: coordinates must lie in [-2^30, 2^30)" % name)
        a = a.astype(np.int32)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)


def dbscan_device(dx: torch.Tensor, dy: torch.Tensor, eps: int, minPts: int, variant: int, cut: int = 0):
    """Labels (int32 CUDA tensor, row order, -1 = noise) + info dict, everything resident in HBM."""
    n = dx.numel()
    labels = torch.empty(n, dtype=torch.int32, device=dx.device)
    info = (C.c_int64 * 8)()
    check(_lib.lib().cloops_dbscan(dx.data_ptr(), dy.data_ptr(), n, int(eps), int(minPts), int(cut), int(variant),
                                   labels.data_ptr(), C.addressof(info), _stream()))
    return labels, dict(zip(INFO_KEYS, (int(v) for v in info)))


def neighbour_counts_device(dx, dy, eps: int, cap: int = 0, cut: int = 0) -> torch.Tensor:
    n = dx.numel()
    out = torch.empty(n, dtype=torch.int32, device=dx.device)
    check(_lib.lib().cloops_neighbour_counts(dx.data_ptr(), dy.data_ptr(), n, int(eps), int(cap), int(cut), out.data_ptr(), _stream()))
    return out


def cluster_summary_device(dx, dy, labels, n_clusters: int, want_row_kind: bool = True):
    """bbox int32[K,4], size int32[K], kind uint8[K], row_kind uint8[n] (all CUDA tensors).  The three
    input arrays may be in any common order (row order or index order)."""
    n = dx.numel()
    k = int(n_clusters)
    dev = dx.device
    bbox = torch.empty((max(k, 1), 4), dtype=torch.int32, device=dev)
    size = torch.empty(max(k, 1), dtype=torch.int32, device=dev)
    kind = torch.empty(max(k, 1), dtype=torch.uint8, device=dev)
    row_kind = torch.empty(max(n, 1), dtype=torch.uint8, device=dev) if want_row_kind else None
    check(_lib.lib().cloops_cluster_summary(dx.data_ptr(), dy.data_ptr(), labels.data_ptr(), n, k, bbox.data_ptr(),
                                            size.data_ptr(), kind.data_ptr(), row_kind.data_ptr() if want_row_kind else None, _stream()))
    return bbox[:k], size[:k], kind[:k], (row_kind[:n] if want_row_kind else None)


def row_kinds_device(labels, kind) -> torch.Tensor:
    n, k = labels.numel(), kind.numel()
    out = torch.empty(max(n, 1), dtype=torch.uint8, device=labels.device)
    check(_lib.lib().cloops_row_kinds(labels.data_ptr(), n, kind.data_ptr(), k, out.data_ptr(), _stream()))
    return out[:n]


class ClusterResult:
    """Output of cluster_and_summarise.  Row-order members (labels, row_kind) are None when the caller
    asked for index order only; index-order members (xs, ys, labels_sorted, kind_sorted) are None for
    blockDBSCAN, which has no strip index."""
    __slots__ = ("labels", "info", "bbox", "size", "kind", "row_kind", "xs", "ys", "labels_sorted", "kind_sorted")


def cluster_and_summarise(dx, dy, eps: int, minPts: int, variant: int, cut: int = 0, rows: bool = True) -> ClusterResult:
    """Cluster one chromosome and reduce the clusters to candidate records.  For v1/v2 the per-cluster
    reduction runs in index order (spatially coherent: few distinct labels per warp) through the resident
    index; ``rows=False`` skips every row-order output (label scatter, per-row kind)."""
    r = ClusterResult()
    r.xs = r.ys = r.labels_sorted = r.kind_sorted = None
    if variant == _lib.BLOCK:
        r.labels, r.info = dbscan_device(dx, dy, eps, minPts, variant, cut)
        r.bbox, r.size, r.kind, r.row_kind = cluster_summary_device(dx, dy, r.labels, r.info["n_clusters"])
        return r
    ix = Index(dx, dy, eps, cut)
    try:
        r.labels, ls, r.info = ix.dbscan(minPts, variant, want_sorted=True, want_rows=rows)
        if ix.n_active:
            r.xs, r.ys = ix.coords()
            r.bbox, r.size, r.kind, _ = cluster_summary_device(r.xs, r.ys, ls, r.info["n_clusters"], want_row_kind=False)
        else:
            r.xs = r.ys = torch.zeros(0, dtype=torch.int32, device=dx.device)
            r.bbox, r.size, r.kind, _ = cluster_summary_device(dx, dy, ls, 0, want_row_kind=False)
        r.labels_sorted = ls
        r.kind_sorted = row_kinds_device(ls, r.kind)
        r.row_kind = row_kinds_device(r.labels, r.kind) if rows else None
    finally:
        ix.close()
    return r


class Index:
    """Resident (strip,u)-sorted index of one chromosome for one eps (cloops_index_* in the C ABI)."""

    def __init__(self, dx: torch.Tensor, dy: torch.Tensor, eps: int, cut: int = 0):
        self.n = dx.numel()
        self.eps = int(eps)
        self._dx, self._dy = dx, dy           # keep inputs alive
        h = C.c_void_p()
        check(_lib.lib().cloops_index_build(dx.data_ptr(), dy.data_ptr(), self.n, int(eps), int(cut), C.byref(h), _stream()))
        self._h = h
        self.n_active = int(_lib.lib().cloops_index_n_active(h))
        if Profile.on:
            Profile.add_stages()

    def count(self, cap: int, out: torch.Tensor | None = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(max(self.n_active, 1), dtype=torch.int32, device=self._dx.device)
        check(_lib.lib().cloops_index_count(self._h, int(cap), out.data_ptr(), _stream()))
        return out

    def dbscan(self, minPts: int, variant: int, want_sorted: bool = False, want_rows: bool = True):
        """labels (row order, or None) [+ labels in index order] + info dict."""
        dev = self._dx.device
        labels = torch.empty(self.n, dtype=torch.int32, device=dev) if want_rows else None
        ls = torch.empty(max(self.n_active, 1), dtype=torch.int32, device=dev) if want_sorted else None
        info = (C.c_int64 * 8)()
        check(_lib.lib().cloops_index_dbscan(self._h, int(minPts), int(variant), labels.data_ptr() if want_rows else None,
                                             ls.data_ptr() if want_sorted else None, C.addressof(info), _stream()))
        info = dict(zip(INFO_KEYS, (int(v) for v in info)))
        return (labels, ls[:self.n_active], info) if want_sorted else (labels, info)

    def coords(self):
        """(X, Y) of the active PETs in index order."""
        xs = torch.empty(max(self.n_active, 1), dtype=torch.int32, device=self._dx.device)
        ys = torch.empty_like(xs)
        check(_lib.lib().cloops_index_coords(self._h, xs.data_ptr(), ys.data_ptr(), _stream()))
        return xs[:self.n_active], ys[:self.n_active]

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().cloops_index_release(self._h, _stream())       # stream-ordered: no host sync
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Coverage:
    """Resident coverage model of one chromosome (cloops_coverage_* in the C ABI): PETs sorted by X
    and by Y in HBM, the GPU form of the reference's getGenomeCoverage (cLoops/cModel.py:45-57)."""

    def __init__(self, dx: torch.Tensor, dy: torch.Tensor):
        self.n = dx.numel()
        self.device = dx.device
        h = C.c_void_p()
        check(_lib.lib().cloops_coverage_build(dx.data_ptr(), dy.data_ptr(), self.n, C.byref(h), _stream()))
        self._h = h

    def _cand(self, cand) -> torch.Tensor:
        cand = np.ascontiguousarray(np.asarray(cand, dtype=np.int64).reshape(-1, 4))
        if cand.size and (cand.min() < -(1 << 31) or cand.max() >= (1 << 31)):
            raise CloopsError("candidate interval outside int32")
        return torch.from_numpy(cand.astype(np.int32)).to(self.device)

    def range_counts(self, cand) -> np.ndarray:
        """cand [m,4] = iva0, iva1, ivb0, ivb1  ->  int32 [m,123] (ra, rb, rab, na[10], nb[10], C[10x10])."""
        d = self._cand(cand)
        m = d.shape[0]
        out = torch.empty((max(m, 1), 123), dtype=torch.int32, device=self.device)
        check(_lib.lib().cloops_range_counts(self._h, d.data_ptr(), m, out.data_ptr(), _stream()))
        self._profile(d, m, 5, 123)
        Profile.d2h_bytes += m * 123 * 4
        return out[:m].cpu().numpy()

    def _profile(self, d, m, win, nout):
        if Profile.on and m:
            Profile.add_stages()
            work = (C.c_uint64 * 2)()
            check(_lib.lib().cloops_range_work(self._h, d.data_ptr(), m, win, C.addressof(work), _stream()))
            Profile.rc_bytes += 8 * (int(work[0]) + int(work[1])) + 4 * nout * m

    def region_pets(self, cand) -> np.ndarray:
        """cand [m,4] -> int32 [m,3] = ra, rb, rab (getPETsforRegions, cModel.py:72-80)."""
        d = self._cand(cand)
        m = d.shape[0]
        out = torch.empty((max(m, 1), 3), dtype=torch.int32, device=self.device)
        check(_lib.lib().cloops_region_pets(self._h, d.data_ptr(), m, out.data_ptr(), _stream()))
        self._profile(d, m, 0, 3)
        Profile.d2h_bytes += m * 3 * 4
        return out[:m].cpu().numpy()

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().cloops_coverage_release(self._h, _stream())    # stream-ordered: no host sync
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DevView:
    """A device buffer owned by the C library, exposed through the CUDA array interface."""

    def __init__(self, ptr, shape, typestr, owner):
        self._owner = owner            # keeps the pass alive while a tensor aliases its memory
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class Pass:
    """One pass of the whole hot path over one chromosome in ONE C-ABI call (cloops_pass_run /
    cloops_pass_run_host): clustering, candidate records, per-PET inter/self membership, coverage model and
    range counts.  Attributes are zero-copy CUDA tensors over buffers owned by the pass."""

    _WHICH = {"bbox": (0, "<i4"), "size": (1, "<i4"), "kind": (2, "|u1"), "xs": (3, "<i4"), "ys": (4, "<i4"),
              "labels_sorted": (5, "<i4"), "member_kind": (6, "|u1"), "cand": (7, "<i4"), "counts": (8, "<i4")}

    def __init__(self, x, y, eps: int, minPts: int, variant: int = _lib.V2, cut: int = 0, score: bool = True, host: bool = False,
                 stats=None, base: "Index | None" = None):
        """``stats`` = (hist int32 CUDA tensor [ROUND_HIST_BINS + 1], mom float64 CUDA tensor [ROUND_MOM]): the round's distance
        accumulators this chromosome adds to (cloops_pass_run_stats)."""
        require_cuda()
        self._keep = (x, y, stats)
        n = x.numel()
        h = C.c_void_p()
        args = (x.data_ptr(), y.data_ptr(), n, int(eps), int(minPts), int(cut), int(variant), 1 if score else 0)
        if base is not None:
            # ``base``: the chromosome's index for this eps built with cut = 0; the round's index is a compaction of it
            if host or base.eps != int(eps):
                raise CloopsError("the base index must be device-resident and built for the same eps")
            self._keep += (base,)
            check(_lib.lib().cloops_pass_run_base(base._h, x.data_ptr(), y.data_ptr(), n, int(minPts), int(cut), int(variant), 1 if score else 0,
                                                  stats[0].data_ptr() if stats is not None else None,
                                                  stats[1].data_ptr() if stats is not None else None, C.byref(h), _stream()))
        elif stats is not None:
            if host:
                raise CloopsError("round statistics need device-resident coordinates")
            check(_lib.lib().cloops_pass_run_stats(*args, stats[0].data_ptr(), stats[1].data_ptr(), C.byref(h), _stream()))
        else:
            fn = _lib.lib().cloops_pass_run_host if host else _lib.lib().cloops_pass_run
            check(fn(*args, C.byref(h), _stream()))
        self._h = h
        sizes, info = (C.c_int64 * 6)(), (C.c_int64 * 8)()
        check(_lib.lib().cloops_pass_sizes(h, C.addressof(sizes), C.addressof(info)))
        self.n_members, self.n_clusters, self.n_candidates, self.scored, self.n_rows = (int(v) for v in sizes[:5])
        self.info = dict(zip(INFO_KEYS, (int(v) for v in info)))
        self._dev = torch.device("cuda", torch.cuda.current_device())
        if Profile.on:
            Profile.add_stages()
            if variant != _lib.BLOCK:
                Profile.rq_bytes += 12 * self.info["n_active"]

    def _view(self, name: str) -> torch.Tensor:
        which, typestr = self._WHICH[name]
        k, nm, m = self.n_clusters, self.n_members, self.n_candidates
        shape = {"bbox": (k, 4), "size": (k,), "kind": (k,), "xs": (nm,), "ys": (nm,), "labels_sorted": (nm,),
                 "member_kind": (nm,), "cand": (m, 4), "counts": (m, 123)}[name]
        dt = torch.int32 if typestr == "<i4" else torch.uint8
        if 0 in shape:
            return torch.zeros(shape, dtype=dt, device=self._dev)
        ptr = _lib.lib().cloops_pass_device_ptr(self._h, which)
        if not ptr:
            return torch.zeros((0,) + tuple(shape[1:]), dtype=dt, device=self._dev)
        return torch.as_tensor(_DevView(ptr, shape, typestr, self), device=self._dev)

    def __getattr__(self, name):
        if name in Pass._WHICH:
            t = self._view(name)
            self.__dict__[name] = t
            return t
        raise AttributeError(name)

    def fetch(self, h_bbox=None, h_kind=None, h_member_kind=None, h_counts=None) -> None:
        """D2H copies into (pinned) host tensors, one synchronisation."""
        ptr = lambda t: t.data_ptr() if t is not None else None
        check(_lib.lib().cloops_pass_fetch(self._h, ptr(h_bbox), ptr(h_kind), ptr(h_member_kind), ptr(h_counts), _stream()))

    def records(self):
        """Host copies (numpy) of bbox int32[k,4], size int32[k], kind u8[k]; one synchronisation."""
        k = self.n_clusters
        bbox, size, kind = np.empty((k, 4), np.int32), np.empty(k, np.int32), np.empty(k, np.uint8)
        if k:
            check(_lib.lib().cloops_pass_fetch_records(self._h, bbox.ctypes.data, size.ctypes.data, kind.ctypes.data, _stream()))
        Profile.d2h_bytes += k * 21
        return bbox, size, kind

    def close(self):
        if getattr(self, "_h", None):
            for name in Pass._WHICH:
                self.__dict__.pop(name, None)
            _lib.lib().cloops_pass_free(self._h, _stream())
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
