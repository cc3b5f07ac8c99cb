import sys, os
import numpy as np
sys.path.insert(0, ".")
from cloops_b200 import device, synth
mode = sys.argv[1] if len(sys.argv) > 1 else "chr21"
if mode == "chr21":
    d = np.load("tests/golden/chr21_pets.npz")
    X, Y = d["X"], d["Y"]
else:
    X, Y = synth.config2(int(mode))
dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
for it in range(3):
    lab, info = device.dbscan_device(dx, dy, 500, 5, 3)
    print(info, flush=True)
