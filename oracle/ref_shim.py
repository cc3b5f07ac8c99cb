"""In-container loader for the UNMODIFIED reference sources (test infrastructure only).

The reference (cLoops 0.93) is Python 2; this image has Python 3.12 only. This module reads the
reference ``.py`` text from ``/root/reference`` at run time, applies the mechanical py2->py3
substitutions listed in SURVEY.md Appendix B (they re-impose py2 integer floor division and
py2 dict/iterator spellings, nothing else) and ``exec``s the result. Nothing from the reference
is copied into this repository.

It exists for ONE purpose: to generate and re-check the golden vectors under ``tests/golden/``
(``oracle/make_golden.py``) and to pin ``oracle/spec.py``. ``/root/reference`` does not exist on
the GPU box; there the loader falls back to ``oracle/_ref`` (an unmodified copy made by ``oracle/make_ref.py``, git-ignored,
shipped with the gpurun snapshot), which is what ``bench.py --impl reference`` / ``cpu_baseline`` time and what the drop-in
test runs.  The product (``cloops_b200/``) never imports this file.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """CLOOPS_REFERENCE, else the mounted reference tree, else the copy made by oracle/make_ref.py (git-ignored,
    shipped to the GPU box by gpurun)."""
    env = os.environ.get("CLOOPS_REFERENCE")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "cLoops", "cDBSCAN2.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "cLoops", "cDBSCAN2.py"))


def _read(name: str) -> str:
    with open(os.path.join(REF_ROOT, "cLoops", name)) as fh:
        return fh.read()


def _sub(src: str, old: str, new: str, *, count: int | None = None) -> str:
    n = src.count(old)
    if n == 0 or (count is not None and n != count):
        raise RuntimeError("shim substitution %r matched %d times" % (old, n))
    return src.replace(old, new)


def _module(name: str, src: str, extra: dict | None = None) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__file__ = "<reference shim %s>" % name
    if extra:
        mod.__dict__.update(extra)
    sys.modules[name] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


_CACHE: dict = {}


def load() -> types.SimpleNamespace:
    """Return namespace with the reference classes/functions (cDBSCAN, cDBSCAN2, blockDBSCAN,
    cModel, pipe, ests, io) executed from the reference's own source text."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)

    # plotting stack is not installed and not on the path under test (settings.py:12-23)
    for stub in ("matplotlib", "matplotlib.pyplot", "seaborn", "pylab", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf"):
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.use = lambda *a, **k: None
            m.rcParams = {}
            m.set_style = lambda *a, **k: None
            m.color_palette = lambda *a, **k: []
            m.PdfPages = object
            sys.modules[stub] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    pkg = types.ModuleType("cLoops")
    pkg.__path__ = []
    sys.modules["cLoops"] = pkg

    utils = _module("cLoops.utils", _read("utils.py"))
    pkg.utils = utils

    # cDBSCAN.py runs unmodified (cDBSCAN.py:84-85 operands are >= 0).
    v1 = _module("cLoops.cDBSCAN", _read("cDBSCAN.py"))

    s = _read("cDBSCAN2.py")
    s = _sub(s, ".iteritems()", ".items()", count=4)          # cDBSCAN2.py:77,117,188,350
    s = _sub(s, "int(x / self.cw)", "int(x // self.cw)", count=1)  # :69 py2 floor
    s = _sub(s, "int(y / self.cw)", "int(y // self.cw)", count=1)  # :70 py2 floor
    v2 = _module("cLoops.cDBSCAN2", s)

    s = _read("blockDBSCAN.py")
    s = _sub(s, "x = x / len(pids)", "x = x // len(pids)", count=1)  # :136
    s = _sub(s, "y = y / len(pids)", "y = y // len(pids)", count=1)  # :137
    blk = _module("cLoops.blockDBSCAN", s)

    s = _read("io.py")
    s = s[: s.index("def jd2washU(")]                          # io.py:292-348 py2 print statements
    s = _sub(s, 'gzip.open(f, "rb")', 'gzip.open(f, "rt")')   # io.py:81,151
    s = _sub(s, "self.cA = (self.startA + self.endA) / 2", "self.cA = (self.startA + self.endA) // 2", count=1)
    s = _sub(s, "self.cB = (self.startB + self.endB) / 2", "self.cB = (self.startB + self.endB) // 2", count=1)
    s = _sub(s, "data.append(map(int, line))", "data.append(list(map(int, line)))", count=1)
    io = _module("cLoops.io", s)

    s = _read("ests.py")
    ests = _module("cLoops.ests", s.replace("from .utils import cFlush", "from cLoops.utils import cFlush"))

    s = _read("cModel.py")
    s = _sub(s, "xrange", "range", count=3)                    # cModel.py:94,207,216
    s = _sub(s, "ca = sum(iva) / 2", "ca = sum(iva) // 2", count=1)
    s = _sub(s, "cb = sum(ivb) / 2", "cb = sum(ivb) // 2", count=1)
    s = _sub(s, "sa = (iva[1] - iva[0]) / 2", "sa = (iva[1] - iva[0]) // 2", count=1)
    s = _sub(s, "sb = (ivb[1] - ivb[0]) / 2", "sb = (ivb[1] - ivb[0]) // 2", count=1)
    s = _sub(s, "step = (sa + sb) / 2", "step = (sa + sb) // 2", count=1)
    s = _sub(s, "keys = ds.keys()", "keys = list(ds.keys())", count=1)
    cmodel = _module("cLoops.cModel", s)

    settings = _module("cLoops.settings", "")
    cplots = _module("cLoops.cPlots", "def plotFragSize(*a, **k):\n    pass\n\ndef plotIntSelCutFrag(*a, **k):\n    pass\n")

    s = _read("pipe.py")
    s = _sub(s, "d = (r[4] + r[5]) / 2 - (r[1] + r[2]) / 2", "d = (r[4] + r[5]) // 2 - (r[1] + r[2]) // 2", count=1)
    pipe = _module("cLoops.pipe", s)

    ns = types.SimpleNamespace(cDBSCAN=v1.cDBSCAN, cDBSCAN2=v2.cDBSCAN, blockDBSCAN=blk.blockDBSCAN,
                               cModel=cmodel, pipe=pipe, ests=ests, io=io, utils=utils)
    _CACHE["ns"] = ns
    return ns
