"""Multi-rank check (run under torchrun): the chromosome-sharded pipeline must write the same .loop
as a single-rank run.  Usage: torchrun --nproc-per-node N tools/dist_pipe_check.py <workdir>"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from cloops_b200 import dist, pipe, synth

work = sys.argv[1]
dist.init_from_env()
rank = dist.rank()
bedpe = os.path.join(work, "in.bedpe")
if rank == 0:
    os.makedirs(work, exist_ok=True)
    with open(bedpe, "w") as fh:
        for ci, n in enumerate((60000, 45000, 30000, 20000, 12000)):
            X, Y = synth.chromosome(n, 3_000_000 + 500_000 * ci, seed=100 + ci, loop_frac=0.25, sigma=400.0)
            for x, y in zip(X.tolist(), Y.tolist()):
                fh.write("chr%d\t%d\t%d\tchr%d\t%d\t%d\tp\t.\t+\t-\n" % (ci + 1, x, x, ci + 1, y, y))
dist.barrier()
os.chdir(work)
out = "multi" if dist.world() > 1 else "single"
pipe.pipe([bedpe], out, [500, 1000], [5], cpu=dist.world(), tmp=0, hic=0)
dist.barrier()
if rank == 0:
    data = open(out + ".loop", "rb").read()
    print("rank0 wrote %s.loop: %d bytes, %d lines, world=%d" % (out, len(data), data.count(b"\n"), dist.world()))
    other = "single.loop" if out == "multi" else "multi.loop"
    if os.path.exists(other):
        same = open(other, "rb").read() == data
        print("IDENTICAL to %s: %s" % (other, same))
        assert same
