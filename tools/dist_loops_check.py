"""Multi-rank check of the resident pipeline (run under torchrun): chromosomes sharded by LPT, per-round all-reduce of the
cut-off statistics, tables gathered -> the loop table must be IDENTICAL to the single-process run on the same genome.
usage: torchrun --nproc-per-node N tools/dist_loops_check.py [total PETs] [config]"""
import hashlib
import sys

import torch

sys.path.insert(0, ".")
import bench
from cloops_b200 import dist, pipe, synth

pets = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
config = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = bench.CONFIGS[config]
dist.init_from_env()
bench.quiet_logs()
rank, world = dist.rank(), dist.world()
counts = synth.genome_counts(pets)
owner = dist.assign(list(range(23)), weights=counts, nranks=world)
cfs = ["mem:%s-%s.jd" % (n, n) for n in synth.CHROMS]
for ci in range(23):
    if owner[ci] == rank:
        name, X, Y = synth.genome_chrom(pets, config, ci)
        pipe._Resident.register(name, X, Y)
run = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], cfg["hic"], weights=counts)
torch.cuda.synchronize()
multi = run["table"].to_csv(sep="\t", index_label="loopId") if run["table"] is not None else ""
print("rank %d/%d: cut %d, %d chromosomes here, table rows %s, sha %s" % (rank, world, run["cut"], len(run["dataI"]),
      None if run["table"] is None else len(run["table"]), hashlib.sha256(multi.encode()).hexdigest()[:16]), flush=True)
dist.barrier()
if rank == 0 and world > 1:
    saved = dict(dist._state)
    dist._state.update(rank=0, world=1)
    pipe._Resident.clear()
    for ci in range(23):
        name, X, Y = synth.genome_chrom(pets, config, ci)
        pipe._Resident.register(name, X, Y)
    solo = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], cfg["hic"], weights=counts)
    single = solo["table"].to_csv(sep="\t", index_label="loopId") if solo["table"] is not None else ""
    dist._state.update(saved)
    print("single process: cut %d, rows %s; IDENTICAL to the %d-rank table: %s" % (solo["cut"], None if solo["table"] is None else len(solo["table"]), world, single == multi), flush=True)
    assert single == multi and solo["cut"] == run["cut"]
dist.barrier()
dist.shutdown()
