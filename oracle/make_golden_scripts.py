"""Golden outputs of the reference's side scripts that call the same boundary:
scripts/quantifyLoops.py, scripts/deLoops and scripts/jd2saturation, EXECUTED from /root/reference through
oracle/ref_shim.py on the bundled chr21 example.  Run:  python oracle/make_golden_scripts.py   (~2 min)

Mechanical py2->py3 substitutions applied to the script text in memory (nothing is copied into the repo):
  * the trailing ``main()`` call of quantifyLoops.py / deLoops is dropped (the functions are called directly);
  * jd2saturation: ``xrange`` -> ``range``; ``os.path.join(os.path.split(f)[:-1])[0]`` (py2 returns the tuple
    unchanged, so this is the directory of f) -> ``os.path.split(f)[0]``.
Inputs: treatment = all chr21 cis PETs (tests/golden/chr21_pets.npz) with the golden ``-m 1`` loop table;
control (deLoops) = the PETs with even row number, same loop table.  jd2saturation: numpy seed 7, eps 750 and
1000, minPts 6, 2 repeats, step 2.
"""
from __future__ import annotations

import logging
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _script(name: str, subs=(), drop_main=False) -> types.ModuleType:
    with open(os.path.join(ref_shim.REF_ROOT, "scripts", name)) as fh:
        src = fh.read()
    for old, new in subs:
        if old not in src:
            raise RuntimeError("substitution %r does not match %s" % (old, name))
        src = src.replace(old, new)
    if drop_main:
        lines = src.rstrip("\n").split("\n")
        if lines[-1].strip() != "main()":
            raise RuntimeError("%s does not end in main()" % name)
        src = "\n".join(lines[:-1]) + "\n"
    mod = types.ModuleType("ref_script_" + name.replace(".", "_"))
    mod.__file__ = "<reference script %s>" % name
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    log = logging.getLogger("ref_script")
    log.addHandler(logging.NullHandler())
    mod.logger = log
    return mod


def write_inputs(work: str):
    """chr21-chr21.jd of treatment (all PETs) and control (even rows) as the reference's txt2jd writes them."""
    import joblib
    d = np.load(os.path.join(GOLD, "chr21_pets.npz"))
    X, Y = d["X"].astype(np.int64), d["Y"].astype(np.int64)
    mat = np.stack([np.arange(len(X)), X, Y], axis=1)
    da, db = os.path.join(work, "trt"), os.path.join(work, "ctl")
    os.makedirs(da)
    os.makedirs(db)
    joblib.dump(mat, os.path.join(da, "chr21-chr21.jd"))
    joblib.dump(mat[::2], os.path.join(db, "chr21-chr21.jd"))
    return da, db


def main():
    ref_shim.load()
    work = tempfile.mkdtemp(prefix="cloops_scripts_")
    cwd = os.getcwd()
    try:
        os.chdir(work)
        da, db = write_inputs(work)
        loop = os.path.join(GOLD, "chr21_m1.loop")

        q = _script("quantifyLoops.py", drop_main=True)
        ra = q.preDs(loop, da, [], ivac=10, ivbc=11)       # column positions of iva/ivb in a pandas>=0.23 .loop file
        q.quantifyLoops(ra, os.path.join(work, "q"), 0, 1)
        shutil.copy(os.path.join(work, "q_quantLoops.txt"), os.path.join(GOLD, "scripts_quantLoops.txt"))

        d = _script("deLoops", drop_main=True)
        ra, rb = d.preDs(loop, da, [], ivac=10, ivbc=11), d.preDs(loop, db, [], ivac=10, ivbc=11)
        d.callDeLoops(ra, rb, "trt", "ctl", 0, 1)
        shutil.copy(os.path.join(work, "trt.deloop"), os.path.join(GOLD, "scripts_trt.deloop"))
        shutil.copy(os.path.join(work, "ctl.deloop"), os.path.join(GOLD, "scripts_ctl.deloop"))

        for f in sorted(os.listdir(GOLD)):
            if f.startswith("scripts_"):
                print(f, os.path.getsize(os.path.join(GOLD, f)))
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
