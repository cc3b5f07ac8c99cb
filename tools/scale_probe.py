"""Developer probe: Hi-C-like density (config 3/4 shape, one large chromosome)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cloops_b200 import _lib, device, hotpath, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
runs = [(5000, 20), (5000, 50), (10000, 20), (10000, 50), (2500, 30)]
X, Y = synth.chromosome(n, 248_956_422, 20240 + 400, loop_frac=0.06, sigma=1500.0)
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
L = _lib.lib()
for eps, mp in runs:
    for variant in (2, 1, 3):
        L.cloops_set_profiling(1)
        torch.cuda.synchronize()
        t = time.time()
        try:
            lab, info = device.dbscan_device(dx, dy, eps, mp, variant)
        except Exception as e:
            print("FAILED", eps, mp, variant, e, flush=True)
            continue
        torch.cuda.synchronize()
        dt = time.time() - t
        print("eps %d minPts %d variant %d: %.1f ms %s" % (eps, mp, variant, dt * 1e3, json.dumps(info)), flush=True)
        print("   stages:", {k: round(v, 2) for k, v in _lib.stage_times().items()}, flush=True)
    L.cloops_set_profiling(0)
    torch.cuda.synchronize()
    t = time.time()
    r = hotpath.run_device(dx, dy, eps, mp)
    torch.cuda.synchronize()
    print("   full hot path %.1f ms, %d inter candidates, giant cluster size %d" % ((time.time() - t) * 1e3, r.cand.shape[0], int(r.size.max())), flush=True)
