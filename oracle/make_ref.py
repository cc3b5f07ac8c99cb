"""Recipe for ``oracle/_ref/``: a byte-for-byte copy of the reference's Python package (and the four side scripts
that call the hot path), plus its bundled chr21 example file, made from ``/root/reference`` where it lies.  Test / measurement infrastructure only.

``oracle/_ref/`` is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so the copy travels to
the GPU box with the snapshot, where ``/root/reference`` does not exist.  ``oracle/ref_shim.py`` executes these UNMODIFIED
files under Python 3 (SURVEY Appendix B substitutions applied in memory).  Consumers: ``bench.py --impl reference`` and
the ``cpu_baseline`` leg (the reference's own runDBSCAN / getIntSig timed on the box's host cores), and the drop-in test
that runs the reference's own ``pipe()`` with ``pipe.DBSCAN`` swapped for the CUDA class.

    python oracle/make_ref.py          # idempotent; also called by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CLOOPS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGE = ["__init__.py", "cDBSCAN.py", "cDBSCAN2.py", "blockDBSCAN.py", "cModel.py", "pipe.py", "io.py", "ests.py", "utils.py",
           "settings.py", "cPlots.py"]
SCRIPTS = ["jd2saturation", "callStripes", "deLoops", "quantifyLoops.py"]
EXAMPLES = ["GSM1872886_GM12878_CTCF_ChIA-PET_chr21_hg38.bedpe.gz"]      # the reference's only fixture: ingest parity


def make(verbose: bool = False) -> bool:
    """-> True when oracle/_ref is complete (copied now or already there)."""
    if not os.path.isfile(os.path.join(SRC, "cLoops", "cDBSCAN2.py")):
        return os.path.isfile(os.path.join(DST, "cLoops", "cDBSCAN2.py"))
    for sub, names in (("cLoops", PACKAGE), ("scripts", SCRIPTS), ("examples", EXAMPLES)):
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
        for name in names:
            a, b = os.path.join(SRC, sub, name), os.path.join(DST, sub, name)
            if not os.path.isfile(a):
                continue
            if not (os.path.isfile(b) and filecmp.cmp(a, b, shallow=False)):
                shutil.copyfile(a, b)
                if verbose:
                    print("copied", os.path.join(sub, name))
    return True


if __name__ == "__main__":
    ok = make(verbose=True)
    print("oracle/_ref", "ready" if ok else "NOT available (no reference tree at %s)" % SRC)
    sys.exit(0 if ok else 1)
