"""Target for the ncu launch list of ONE default bench step (config 4: 200 M PETs, 12 rounds, scoring), device part only:
   ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file X.csv python tools/ncu_step.py [config] [pets] [nchrom]
nchrom < 23 restricts the step to the first chromosomes (same per-chromosome sizes: ncu serialises every launch, the whole
genome takes a quarter of an hour to capture); one chromosome in flight at a time."""
import sys

import torch

sys.path.insert(0, ".")
import bench
from cloops_b200 import _lib, pipe, synth

config = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = bench.CONFIGS[config]
pets = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["pets"]
nchrom = int(sys.argv[3]) if len(sys.argv) > 3 else 23
bench.quiet_logs()
pipe.STREAMS = 1
counts = synth.genome_counts(pets)[:nchrom]
chroms = bench.generate(list(range(nchrom)), lambda ci: synth.genome_chrom(pets, config, ci))
for name, X, Y in chroms:
    pipe._Resident.register(name, X, Y)
cfs = ["mem:%s-%s.jd" % (n, n) for n in synth.CHROMS[:nchrom]]
l0 = _lib.lib().cloops_kernel_launches()
r = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], cfg["hic"], weights=counts, tail=False)
torch.cuda.synchronize()
print("launches", _lib.lib().cloops_kernel_launches() - l0, "cut", r["cut"])
