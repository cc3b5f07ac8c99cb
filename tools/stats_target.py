"""ncu target: the distance-statistics kernels of one config-4 chr1 pass (cut 11500) and the round's order statistics.
   ncu --set full --clock-control none --import-source on -k regex:'dist_stats|hist_' -o gpurun_out/stats python tools/stats_target.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from cloops_b200 import _lib, device, synth

name, X, Y = synth.genome_chrom(200_000_000, 4, 0)
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
hist = torch.zeros(_lib.ROUND_HIST_BINS + 1, dtype=torch.int32, device="cuda")
mom = torch.zeros(_lib.ROUND_MOM, dtype=torch.float64, device="cuda")
for rep in range(2):
    p = device.Pass(dx, dy, 5000, 20, _lib.V2, 11500, score=False, stats=(hist, mom))
    p.records()
    p.close()
mid = np.zeros(2, np.int64)
hm = np.zeros(_lib.ROUND_MOM, np.float64)
_lib.check(_lib.lib().cloops_round_middle(hist.data_ptr(), mom.data_ptr(), mid.ctypes.data, hm.ctypes.data, torch.cuda.current_stream().cuda_stream))
print(name, len(X), mid, hm[:9])
