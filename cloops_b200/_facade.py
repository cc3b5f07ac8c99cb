"""Shared implementation of the three drop-in clusterer classes.

Call surface kept from the reference (cLoops/pipe.py:70-71, scripts/jd2saturation:69-70,
scripts/callStripes:51-52): ``DBSCAN(mat, eps, minPts)`` does all the work in the constructor and the
caller reads ``.labels`` -- a ``{pointId: clusterId}`` dict from which noise points are absent.
Additions: ``.labels_array`` (int32, row order, -1 = noise) so large callers can skip the dict, and
``.info``.  ``mat`` is never modified.
"""
from __future__ import annotations

import numpy as np

from . import device
from ._lib import CloopsError


def _close_gaps(c, eps: int):
    """Order-preserving re-coding of one coordinate axis: differences <= eps are kept exactly, larger gaps between
    consecutive distinct values become eps + 1, the smallest value becomes 0."""
    c = np.asarray(c, dtype=np.int64)
    order = np.argsort(c, kind="stable")
    s = c[order]
    out = np.empty_like(c)
    out[order] = np.concatenate([[0], np.cumsum(np.minimum(np.diff(s), eps + 1))])
    return out


class _GpuDBSCAN:
    _variant = None

    def __init__(self, mat, eps, minPts):
        self.eps = eps
        self.minPts = minPts
        mat = np.asarray(mat)
        if mat.ndim != 2 or (mat.shape[0] and mat.shape[1] < 3):
            raise CloopsError("mat must be an [N,3] array of [pointId, X, Y] rows")
        if int(eps) != eps or int(minPts) != minPts:
            raise CloopsError("eps and minPts must be integers (PET coordinates are integer base pairs)")
        n = mat.shape[0]
        self._ids = mat[:, 0] if n else np.zeros(0, np.int64)
        self._labels = None
        if n == 0:
            # cDBSCAN2 returns {} on empty input (cDBSCAN2.py:114-192); v1/block raise IndexError
            # (cDBSCAN.py:77, blockDBSCAN.py:74) -- kept.
            if self._variant != 2:
                raise IndexError("index 0 is out of bounds for axis 0 with size 0")
            self.labels_array = np.zeros(0, np.int32)
            self.info = {}
            return
        X, Y = mat[:, 1], mat[:, 2]
        if self._variant == 1 and n and (max(abs(int(X.min())), abs(int(X.max())), abs(int(Y.min())), abs(int(Y.max()))) >= device.COORD_LIMIT):
            # scripts/callStripes:42-43 scales one axis by 50 before clustering with the v1 class: coordinates beyond int32.
            # v1's labels depend on the Manhattan neighbour relation and on row order only (its grid offset is
            # label-neutral), and that relation survives shrinking every gap between consecutive distinct values of an axis
            # that is wider than eps down to eps + 1 -- so sparse wide-range inputs are folded back into int32.
            X, Y = _close_gaps(X, int(eps)), _close_gaps(Y, int(eps))
        dx = device.to_device_i32(X, "X")
        dy = device.to_device_i32(Y, "Y")
        lab, self.info = device.dbscan_device(dx, dy, int(eps), int(minPts), self._variant)
        self.labels_array = lab.cpu().numpy()

    @property
    def labels(self) -> dict:
        if self._labels is None:
            keep = self.labels_array >= 0
            self._labels = dict(zip(self._ids[keep].tolist(), self.labels_array[keep].tolist()))
        return self._labels
