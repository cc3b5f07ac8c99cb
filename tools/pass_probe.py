"""Stage times of single passes at the shapes of config 4 / 3 / 2 (run on the GPU box):
   python tools/pass_probe.py [chr index] -> per (eps, minPts, cut): stage table, launches."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cloops_b200 import _lib, device, synth

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 0
name, X, Y = synth.genome_chrom(200_000_000, 4, ci)
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
L = _lib.lib()
hist = torch.zeros(_lib.ROUND_HIST_BINS + 1, dtype=torch.int32, device="cuda")
mom = torch.zeros(_lib.ROUND_MOM, dtype=torch.float64, device="cuda")
cases = [(5000, 50, 0), (5000, 20, 11500), (7500, 40, 11500), (10000, 20, 11500)]
for eps, mp, cut in cases:
    for rep in range(3):
        L.cloops_set_profiling(1 if rep == 2 else 0)
        torch.cuda.synchronize()
        l0 = L.cloops_kernel_launches()
        t0 = time.perf_counter()
        p = device.Pass(dx, dy, eps, mp, _lib.V2, cut, score=False, stats=(hist, mom))
        b, s, k = p.records()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = _lib.stage_times()
        info = p.info
        p.close()
    print("%s n=%d eps=%d mp=%d cut=%d: wall %.2f ms, stages %.2f ms, launches %d, n_act %d clusters %d dead %d" %
          (name, len(X), eps, mp, cut, dt * 1e3, sum(st.values()), L.cloops_kernel_launches() - l0, info["n_active"], info["n_clusters"], info["n_dead"]))
    print("   " + "  ".join("%s=%.2f" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])))
L.cloops_set_profiling(0)
# scoring shapes: candidates of the last pass
p = device.Pass(dx, dy, 10000, 20, _lib.V2, 11500, score=False)
b, s, k = p.records()
p.close()
cand = b[k == 1].astype(np.int64)
cand[:, 0] = np.maximum(cand[:, 0], 0)
cand[:, 2] = np.maximum(cand[:, 2], 0)
cov = device.Coverage(dx, dy)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r3 = cov.region_pets(cand)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    keep = np.flatnonzero(r3[:, 2] >= 50)
    rc = cov.range_counts(cand[keep])
    torch.cuda.synchronize(); t2 = time.perf_counter()
print("scoring: %d candidates region_pets %.2f ms; %d kept range_counts %.2f ms; mean ra %.0f rb %.0f" %
      (len(cand), (t1 - t0) * 1e3, len(keep), (t2 - t1) * 1e3, r3[:, 0].mean(), r3[:, 1].mean()))
