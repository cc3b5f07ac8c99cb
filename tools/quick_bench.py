"""Developer probe: stage timings of the clustering path on the config-2 synthetic set."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cloops_b200 import _lib, device, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
eps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
mp = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t = time.time()
X, Y = synth.config2(n)
print("synth %.1fs" % (time.time() - t), flush=True)
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
L = _lib.lib()
for variant in (2, 1, 3):
    try:
        for it in range(3):
            L.cloops_set_profiling(1 if it == 2 else 0)
            torch.cuda.synchronize()
            t = time.time()
            lab, info = device.dbscan_device(dx, dy, eps, mp, variant)
            torch.cuda.synchronize()
            dt = time.time() - t
        print("variant", variant, "wall %.2f ms" % (dt * 1e3), json.dumps(info), flush=True)
        print("  stages:", {k: round(v, 3) for k, v in _lib.stage_times().items()}, flush=True)
    except Exception as e:
        print("variant", variant, "failed:", e)
L.cloops_set_profiling(0)
ix = device.Index(dx, dy, eps)
out = torch.empty(ix.n_active, dtype=torch.int32, device=dx.device)
for cap in (mp, 0):
    for _ in range(3):
        ix.count(cap, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ix.count(cap, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("count kernel cap=%d: %.3f ms  -> %.1f GB/s algorithmic (12 B/PET)" % (cap, ms, 12 * ix.n_active / ms / 1e6))
print("mean count", float(out.float().mean()), "max", int(out.max()))
