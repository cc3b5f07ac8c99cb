"""ctypes binding of the C ABI in include/cloops_b200.h.  There is no fallback: if the shared library
is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcloops_b200.so")

V1, V2, BLOCK = 1, 2, 3
VARIANTS = {"v1": V1, "v2": V2, "block": BLOCK}
ROUND_HIST_BINS, ROUND_MOM = 1 << 20, 16


class CloopsError(RuntimeError):
    pass


_lib = None

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

_SIGNATURES = {
    "cloops_last_error": (C.c_char_p, []),
    "cloops_version": (C.c_char_p, []),
    "cloops_kernel_launches": (_i64, []),
    "cloops_set_profiling": (None, [C.c_int]),
    "cloops_stage_count": (C.c_int, []),
    "cloops_stage_name": (C.c_char_p, [C.c_int]),
    "cloops_stage_ms": (C.c_float, [C.c_int]),
    "cloops_workspace_release": (C.c_int, []),
    "cloops_dbscan": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "cloops_dbscan_host": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "cloops_neighbour_counts": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "cloops_index_build": (C.c_int, [_vp, _vp, _i64, _i32, _i32, C.POINTER(_vp), _vp]),
    "cloops_index_free": (None, [_vp]),
    "cloops_index_release": (None, [_vp, _vp]),
    "cloops_index_n_active": (_i64, [_vp]),
    "cloops_index_count": (C.c_int, [_vp, _i32, _vp, _vp]),
    "cloops_index_dbscan": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "cloops_index_coords": (C.c_int, [_vp, _vp, _vp, _vp]),
    "cloops_row_kinds": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "cloops_cluster_summary": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "cloops_coverage_build": (C.c_int, [_vp, _vp, _i64, C.POINTER(_vp), _vp]),
    "cloops_coverage_free": (None, [_vp]),
    "cloops_coverage_release": (None, [_vp, _vp]),
    "cloops_range_counts": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "cloops_range_work": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "cloops_region_pets": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "cloops_remove_dup": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "cloops_combine_rounds": (C.c_int, [_vp, _vp, _i64, _vp]),
    "cloops_bedpe_parse": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i64, C.c_int, C.POINTER(_vp)]),
    "cloops_bedpe_lines": (_i64, [_vp]),
    "cloops_bedpe_bare_cr": (_i64, [_vp]),
    "cloops_bedpe_n_chroms": (C.c_int, [_vp]),
    "cloops_bedpe_chrom": (_vp, [_vp, C.c_int, _vp, _vp]),
    "cloops_bedpe_fetch": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "cloops_bedpe_n_odd": (_i64, [_vp]),
    "cloops_bedpe_odd": (_vp, [_vp, _i64, _vp, _vp]),
    "cloops_bedpe_free": (None, [_vp]),
    "cloops_pass_run": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, C.POINTER(_vp), _vp]),
    "cloops_pass_run_stats": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, C.POINTER(_vp), _vp]),
    "cloops_pass_run_base": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, C.POINTER(_vp), _vp]),
    "cloops_round_middle": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "cloops_pass_run_host": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, C.POINTER(_vp), _vp]),
    "cloops_pass_sizes": (C.c_int, [_vp, _vp, _vp]),
    "cloops_pass_device_ptr": (_vp, [_vp, C.c_int]),
    "cloops_pass_fetch": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "cloops_pass_fetch_records": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "cloops_pass_free": (None, [_vp, _vp]),
}

EXPORTS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise CloopsError(
                "libcloops_b200.so is not built (%s). Run `python -m cloops_b200._build`; there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise CloopsError("libcloops_b200 error %d: %s" % (rc, lib().cloops_last_error().decode()))


def stage_times() -> dict:
    L = lib()
    out = {}
    for i in range(L.cloops_stage_count()):
        name = L.cloops_stage_name(i).decode()
        out[name] = out.get(name, 0.0) + float(L.cloops_stage_ms(i))
    return out
