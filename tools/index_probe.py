"""Developer probe: stage times of the (strip,u) index build on the config-2 set (or one build per run under ncu)."""
import sys

import torch

sys.path.insert(0, ".")
from cloops_b200 import _lib, device, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
eps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
X, Y = synth.config2(n)
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L = _lib.lib()
L.cloops_set_profiling(1)
acc = {}
for it in range(6):
    flush.fill_(1)
    ix = device.Index(dx, dy, eps)
    if it > 0:
        for k, v in _lib.stage_times().items():
            acc[k] = acc.get(k, 0.0) + v / 5
    ix.close()
torch.cuda.synchronize()
print({k: round(v, 4) for k, v in acc.items()}, "total %.4f ms" % sum(acc.values()))
