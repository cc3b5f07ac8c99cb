"""cProfile of one resident config-4 step (host-side view): python tools/step_profile.py [pets]"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
from cloops_b200 import pipe, synth

pets = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
cfg = bench.CONFIGS[4]
bench.quiet_logs()
counts = synth.genome_counts(pets)
chroms = bench.generate(list(range(23)), lambda ci: synth.genome_chrom(pets, 4, ci))
for name, X, Y in chroms:
    pipe._Resident.register(name, X, Y)
cfs = ["mem:%s-%s.jd" % (n, n) for n in synth.CHROMS]
for tail in (False, True):
    pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], 1, weights=counts, tail=tail)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pr = cProfile.Profile()
    pr.enable()
    r = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], 1, weights=counts, tail=tail)
    torch.cuda.synchronize()
    pr.disable()
    print("tail=%s wall %.1f ms" % (tail, (time.perf_counter() - t0) * 1e3))
    pstats.Stats(pr, stream=sys.stdout).sort_stats("tottime").print_stats(22)
