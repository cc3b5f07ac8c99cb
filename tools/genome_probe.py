"""Developer probe: the whole multi-round pipeline (cLoops -m 4 style) on a synthetic 23-chromosome
HiChIP-like genome, entering at the .jd level; prints where the wall time goes."""
import cProfile
import io
import os
import pstats
import shutil
import sys
import tempfile
import time

import joblib
import numpy as np

sys.path.insert(0, ".")
from cloops_b200 import pipe, synth

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
work = tempfile.mkdtemp(prefix="genome_probe_")
t = time.time()
fs = []
for name, X, Y in synth.genome(n_total, config=3):
    f = os.path.join(work, "%s-%s.jd" % (name, name))
    joblib.dump(np.stack([np.arange(len(X)), X, Y], axis=1).astype(np.int64), f)
    fs.append(f)
print("generated %d PETs in %d chromosomes: %.1f s" % (n_total, len(fs), time.time() - t), flush=True)
eps, minPts = [2500, 5000, 7500, 10000], [30, 20]
pr = cProfile.Profile()
pr.enable()
t0 = time.time()
dataI, cut, cuts = {}, 0, [0]
for ep in eps:
    for m in minPts:
        t = time.time()
        dataI_2, dataS_2, n_dis, n_dss, cut_2 = pipe._round(fs, ep, m, cut)
        if cut_2 is not None:
            cuts.append(cut_2)
            cut = cut_2
        dataI = pipe.combineTwice(dataI, dataI_2)
        print("round eps %d minPts %d: %.2f s, cut -> %d, candidates so far %d" %
              (ep, m, time.time() - t, cut, sum(len(v["records"]) for v in dataI.values())), flush=True)
t1 = time.time()
cut = min(c for c in cuts if c > 0)
dataI = pipe.filterClusterByDis(dataI, cut)
ncand = sum(len(v["records"]) for v in dataI.values())
print("clustering rounds: %.2f s ; %d candidates after distance filter" % (t1 - t0, ncand), flush=True)
so = sys.stdout
sys.stdout = io.StringIO()
try:
    e = pipe.runStat(dataI, minPts, 0, 1, os.path.join(work, "out"), 1)
finally:
    sys.stdout = so
t2 = time.time()
pr.disable()
print("scoring (GPU range counts + host tail): %.2f s" % (t2 - t1))
n_loops = sum(1 for _ in open(os.path.join(work, "out.loop"))) - 1 if os.path.exists(os.path.join(work, "out.loop")) else 0
print("loops: %d ; total %.2f s = %.0f PETs/s end to end" % (n_loops, t2 - t0, n_total / (t2 - t0)))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
print("\n".join(s.getvalue().splitlines()[:45]))
shutil.rmtree(work, ignore_errors=True)
