"""Config 5 (eps x minPts sweep on one 50 M-PET chromosome): time of every index build and clustering, printed as it goes."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
from cloops_b200 import _lib, device, synth

cfg = bench.CONFIGS[5]
pets = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["pets"]
X, Y = synth.chromosome(pets, bench.CHROM_LEN_SINGLE, 20240 + 500, loop_frac=0.08, sigma=500.0)
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
L = _lib.lib()
L.cloops_set_profiling(1)
for ep in cfg["eps"]:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ix = device.Index(dx, dy, ep)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("eps %d: index %.1f ms" % (ep, (t1 - t0) * 1e3), flush=True)
    for m in cfg["minPts"]:
        t0 = time.perf_counter()
        _, ls, info = ix.dbscan(m, _lib.V2, want_sorted=True, want_rows=False)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        st = _lib.stage_times()
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print("   minPts %d: %.1f ms  clusters %d core %d  %s" % (m, (t1 - t0) * 1e3, info["n_clusters"], info["n_core"],
              " ".join("%s=%.1f" % kv for kv in top)), flush=True)
    ix.close()
