"""Developer probe: per-launch time of the region-query kernel (CUDA events, L2 flushed between launches)
on the synthetic shapes of BASELINE.json, with the counts checked against a second launch.
usage: python tools/count_probe.py [ncu]   ("ncu": one launch on config 2, for `ncu -k regex:count_kernel`)"""
import sys

import torch

sys.path.insert(0, ".")
from cloops_b200 import device, synth

ncu_mode = len(sys.argv) > 1 and sys.argv[1] == "ncu"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(ix, cap, out, iters=10):
    ts = []
    for _ in range(3):
        ix.count(cap, out)
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.count(cap, out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def run(name, X, Y, eps, caps, cut=0):
    dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
    ix = device.Index(dx, dy, eps, cut)
    n = ix.n_active
    for cap in caps:
        out = torch.full((n,), -7, dtype=torch.int32, device="cuda")
        if ncu_mode:
            ix.count(cap, out)
            torch.cuda.synchronize()
            continue
        med, best = timed(ix, cap, out)
        print("%s eps=%d cap=%d: median %.1f us best %.1f us -> %.0f GB/s (12 B/PET)  mean count %.3f" %
              (name, eps, cap, med * 1e3, best * 1e3, 12 * n / med / 1e6, float(out.float().mean())), flush=True)
    ix.close()


X, Y = synth.config2(10_000_000)
run("config2-10M", X, Y, 1000, [5] if ncu_mode else [5, 0, 3, 9, 12])
if ncu_mode:
    X, Y = synth.genome_chrom(200_000_000, 4, 0)[1:]
    run("config4-chr1", X, Y, 5000, [20], cut=11500)        # a later round of -m 3: the cut has removed the diagonal
    run("config4-chr1", X, Y, 5000, [50])                   # round 1
if not ncu_mode:
    X, Y = synth.genome_chrom(200_000_000, 4, 0)[1:]
    run("config4-chr1-cut11500", X, Y, 5000, [20], cut=11500)
    run("config4-chr1-cut11500", X, Y, 10000, [50], cut=11500)
    run("config4-chr1", X, Y, 5000, [50])
    run("config2-10M", X, Y, 250, [5])
    run("config2-10M", X, Y, 4000, [5])
    X, Y = synth.chromosome(16_000_000, 248_956_422, 20240 + 400, loop_frac=0.06, sigma=1500.0)
    for eps, mp in ((5000, 20), (10000, 50), (2500, 30)):
        run("hic-16M", X, Y, eps, [mp])
    X, Y = synth.chromosome(50_000_000, 248_956_422, 20240 + 500, loop_frac=0.08, sigma=500.0)
    run("config5-50M", X, Y, 1000, [5])
    run("config5-50M", X, Y, 5000, [20])
