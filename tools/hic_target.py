"""Short target for ncu captures at Hi-C density: v2 rounds on a 16 M-PET chr1-sized chromosome."""
import sys

import torch

sys.path.insert(0, ".")
from cloops_b200 import _lib, device, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
eps = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
mp = int(sys.argv[3]) if len(sys.argv) > 3 else 20
X, Y = synth.chromosome(n, 248_956_422, 20240 + 400, loop_frac=0.06, sigma=1500.0)
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
for _ in range(2):
    lab, info = device.dbscan_device(dx, dy, eps, mp, _lib.V2)
torch.cuda.synchronize()
print(info)
