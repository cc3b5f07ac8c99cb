"""Short target for ncu captures: a few passes of the clustering path on the config-2 set."""
import sys

import torch

sys.path.insert(0, ".")
from cloops_b200 import _lib, device, hotpath, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
X, Y = synth.config2(n)
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
for _ in range(passes):
    hotpath.run_device(dx, dy, 1000, 5)
torch.cuda.synchronize()
print("launches", _lib.lib().cloops_kernel_launches())
