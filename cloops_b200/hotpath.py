"""One pass of the whole hot path for one chromosome: cluster -> candidate records -> permuted-
background range counts of every inter-ligation candidate.  This is the unit ``bench.py`` times and
the composition ``pipe.singleDBSCAN`` + ``cModel.getIntSig`` perform per chromosome."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, device
from ._lib import check


_SIDE = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


class HotPathResult:
    """labels / row_kind: row order (None unless rows=True); kind_sorted: per-PET kind in index order;
    bbox, size, kind: per cluster id; cand, counts: per inter-ligation candidate (ascending id)."""
    __slots__ = ("labels", "info", "bbox", "size", "kind", "row_kind", "kind_sorted", "cand", "counts")


def run_device_separate(dx: torch.Tensor, dy: torch.Tensor, eps: int, minPts: int, variant: int = _lib.V2, cut: int = 0,
                        score: bool = True, rows: bool = False) -> HotPathResult:
    """Inputs and outputs resident in HBM.  What the reference's per-chromosome step returns
    (cLoops/pipe.py:52-110: candidate records and the membership of dis / dss) plus the range counts of
    every inter-ligation candidate, int32 [n_inter, 123] in ascending cluster-id order (pipe.py:97,
    cModel.py:281-295).  Cluster labels in row order are produced only on request (``rows=True``)."""
    # The coverage model (two radix sorts) does not depend on the clustering: build it on a side stream
    # while the latency-bound clustering kernels (union-find, border passes) run on the main stream.
    main = torch.cuda.current_stream()
    cov = None
    if score:
        side = _side_stream(dx.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            cov = device.Coverage(dx, dy)
    c = device.cluster_and_summarise(dx, dy, eps, minPts, variant, cut, rows=rows or variant == _lib.BLOCK)
    r = HotPathResult()
    r.labels, r.info, r.bbox, r.size, r.kind, r.row_kind = c.labels, c.info, c.bbox, c.size, c.kind, c.row_kind
    r.kind_sorted = c.kind_sorted if c.kind_sorted is not None else c.row_kind
    r.cand = r.counts = None
    if score:
        cand = r.bbox[r.kind == 1].clamp_(min=0)[:, [0, 1, 2, 3]].contiguous()     # max(0, .) of cModel.py:281-282
        m = cand.shape[0]
        main.wait_stream(side)
        out = torch.empty((max(m, 1), 123), dtype=torch.int32, device=dx.device)
        if m:
            check(_lib.lib().cloops_range_counts(cov._h, cand.data_ptr(), m, out.data_ptr(), main.cuda_stream))
        cov.close()
        r.cand, r.counts = cand, out[:m]
    return r


def run_device(dx: torch.Tensor, dy: torch.Tensor, eps: int, minPts: int, variant: int = _lib.V2, cut: int = 0,
               score: bool = True) -> "device.Pass":
    """The same pass through the single C-ABI entry point ``cloops_pass_run``; returns the resident
    ``device.Pass`` (bbox, size, kind, cand, counts, xs, ys, labels_sorted, member_kind, info)."""
    return device.Pass(dx, dy, eps, minPts, variant, cut, score=score)


class HostStep:
    """The same pass from HOST buffers: pinned int32 X, Y in; candidate records, per-PET kind bytes and
    range counts out, through the host entry points of the C ABI.  Pinned result buffers are allocated
    once and re-used; one synchronisation at the end."""

    def __init__(self, n: int, device_index: int | None = None):
        self._pinned = {}
        self.h2d_bytes = 2 * 4 * n
        self.d2h_bytes = 0

    def __call__(self, hx: torch.Tensor, hy: torch.Tensor, eps: int, minPts: int, variant: int = _lib.V2):
        """cloops_pass_run_host (H2D of the pinned coordinates + the whole pass) and cloops_pass_fetch
        (D2H of records, per-PET membership and range counts into pinned buffers, one synchronisation)."""
        p = device.Pass(hx, hy, eps, minPts, variant, 0, score=True, host=True)
        bbox = self._buf("bbox", (p.n_clusters, 4), torch.int32)
        kind = self._buf("kind", (p.n_clusters,), torch.uint8)
        members = self._buf("members", (p.n_members,), torch.uint8)
        counts = self._buf("counts", (p.n_candidates, 123), torch.int32)
        p.fetch(bbox, kind, members, counts)
        self.d2h_bytes = members.numel() + bbox.numel() * 4 + kind.numel() + counts.numel() * 4
        self.info = p.info
        p.close()
        return bbox, kind, counts, members

    def _buf(self, name: str, shape, dtype) -> torch.Tensor:
        need = 1
        for d in shape:
            need *= d
        buf = self._pinned.get(name)
        if buf is None or buf.numel() < need or buf.dtype != dtype:
            buf = torch.empty(max(need * 2, 1024), dtype=dtype).pin_memory()
            self._pinned[name] = buf
        return buf[:need].view(shape)
