"""ctypes binding of oracle/coracle.c (the C restatement of the hot path).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this module; cloops_b200/ never does.

``build()`` compiles ``oracle/coracle.c`` with gcc into ``oracle/_build/libcoracle.so`` (git-ignored; shipped to the GPU box
by gpurun, rebuilt there if missing -- gcc is in the image)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "coracle.c")
LIB = os.path.join(HERE, "_build", "libcoracle.so")
V1, V2, BLOCK = 1, 2, 3
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        p = C.c_void_p
        L.coracle_counts.argtypes = [p, p, C.c_int64, C.c_int64, C.c_int64, p]
        L.coracle_dbscan.argtypes = [p, p, C.c_int64, C.c_int64, C.c_int64, C.c_int, p, p]
        L.coracle_records.argtypes = [p, p, p, C.c_int64, C.c_int64, p, p, p]
        L.coracle_range_counts.argtypes = [p, p, C.c_int64, p, C.c_int64, p]
        for f in (L.coracle_counts, L.coracle_dbscan, L.coracle_records, L.coracle_range_counts):
            f.restype = C.c_int
        L.coracle_version.restype = C.c_char_p
        _lib = L
    return _lib


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def neighbour_counts(X, Y, eps: int, cap: int = 0) -> np.ndarray:
    """min(n(p), cap) per row (cap <= 0: exact), n(p) = #{q: |dX|+|dY| <= eps} including p."""
    X, Y = _i64(X), _i64(Y)
    out = np.empty(len(X), np.int32)
    rc = lib().coracle_counts(X.ctypes.data, Y.ctypes.data, len(X), int(eps), int(cap), out.ctypes.data)
    if rc:
        raise RuntimeError("coracle_counts failed (%d)" % rc)
    return out


def dbscan(X, Y, eps: int, minPts: int, variant: int = V2, return_info: bool = False):
    """Labels in row order (-1 = absent from the reference's labels dict) of cDBSCAN (V1), cDBSCAN2 (V2), blockDBSCAN (BLOCK)."""
    X, Y = _i64(X), _i64(Y)
    out = np.empty(len(X), np.int32)
    info = np.zeros(8, np.int64)
    rc = lib().coracle_dbscan(X.ctypes.data, Y.ctypes.data, len(X), int(eps), int(minPts), int(variant), out.ctypes.data, info.ctypes.data)
    if rc:
        raise RuntimeError("coracle_dbscan failed (%d)" % rc)
    if return_info:
        return out, {"clusters": int(info[0]), "components": int(info[1]), "core": int(info[2]), "dead": int(info[3])}
    return out


def cluster_records(X, Y, labels):
    """-> bbox int64 [K,4] (minX,maxX,minY,maxY), size [K], kind u8 [K] (0 dropped / 1 inter / 2 self), K = max label + 1."""
    X, Y = _i64(X), _i64(Y)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    k = int(labels.max()) + 1 if len(labels) else 0
    k = max(k, 0)
    bbox = np.zeros((k, 4), np.int64)
    size = np.zeros(k, np.int64)
    kind = np.zeros(k, np.uint8)
    rc = lib().coracle_records(X.ctypes.data, Y.ctypes.data, labels.ctypes.data, len(X), k, bbox.ctypes.data, size.ctypes.data, kind.ctypes.data)
    if rc:
        raise RuntimeError("coracle_records failed (%d)" % rc)
    return bbox, size, kind


def range_counts(X, Y, cand) -> np.ndarray:
    """cand int [M,4] = (iva0, iva1, ivb0, ivb1) as getMultiplePsFdr receives them (already clamped at 0,
    cModel.py:281-282) -> int64 [M,123]."""
    X, Y = _i64(X), _i64(Y)
    cand = _i64(cand).reshape(-1, 4)
    out = np.zeros((len(cand), 123), np.int64)
    rc = lib().coracle_range_counts(X.ctypes.data, Y.ctypes.data, len(X), cand.ctypes.data, len(cand), out.ctypes.data)
    if rc:
        raise RuntimeError("coracle_range_counts failed (%d)" % rc)
    return out


def hot_path(X, Y, eps: int, minPts: int):
    """One pass of the whole hot path (v2 -> candidate records -> range counts of every inter-ligation candidate);
    -> (labels, inter bbox [K,4], counts [K,123]) like oracle.spec.hot_path_cpu."""
    lab = dbscan(X, Y, eps, minPts, V2)
    bbox, size, kind = cluster_records(X, Y, lab)
    inter = bbox[kind == 1]
    cand = inter.copy()
    cand[:, 0] = np.maximum(cand[:, 0], 0)
    cand[:, 2] = np.maximum(cand[:, 2], 0)
    return lab, inter, range_counts(X, Y, cand)
