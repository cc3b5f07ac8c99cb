"""CPU-side checks of the C-ABI library: it loads and exports every symbol include/cloops_b200.h
declares (no compute calls: there is no GPU in the build container)."""
import os
import re

from cloops_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "cloops_b200.h")).read()
    declared = set(re.findall(r"CLOOPS_API[^;(]*?\b(cloops_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.EXPORTS)
    assert b"sm_100a" in L.cloops_version()
    assert L.cloops_kernel_launches() == 0


def test_no_oracle_in_product():
    """The product must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "cloops_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(dp, f)).read()), f
