"""Target for the ncu launch list of ONE default bench step (config 4: 200 M PETs, 12 rounds, scoring), device part only:
   ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file X.csv python tools/ncu_step.py"""
import sys

import torch

sys.path.insert(0, ".")
import bench
from cloops_b200 import _lib, pipe, synth

config = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = bench.CONFIGS[config]
pets = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["pets"]
bench.quiet_logs()
counts = synth.genome_counts(pets)
chroms = bench.generate(list(range(23)), lambda ci: synth.genome_chrom(pets, config, ci))
for name, X, Y in chroms:
    pipe._Resident.register(name, X, Y)
cfs = ["mem:%s-%s.jd" % (n, n) for n in synth.CHROMS]
l0 = _lib.lib().cloops_kernel_launches()
r = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], cfg["hic"], weights=counts, tail=False)
torch.cuda.synchronize()
print("launches", _lib.lib().cloops_kernel_launches() - l0, "cut", r["cut"])
