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
        dx = device.to_device_i32(mat[:, 1], "X")
        dy = device.to_device_i32(mat[:, 2], "Y")
        lab, self.info = device.dbscan_device(dx, dy, int(eps), int(minPts), self._variant)
        self.labels_array = lab.cpu().numpy()

    @property
    def labels(self) -> dict:
        if self._labels is None:
            keep = self.labels_array >= 0
            self._labels = dict(zip(self._ids[keep].tolist(), self.labels_array[keep].tolist()))
        return self._labels
