"""Workload for ncu captures (GPU box): a few resident-list E+F steps of a water box with the chosen pair-kernel variant.
    ncu --set full --import-source on -k regex:<kernels> -c <n> python tools/ncu_target.py [n_mol] [pair_block]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from apax_b200 import runtime as rt, synthetic as syn

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 333334
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pos, Z, cell = syn.water_box(n_mol, 0)
n = len(Z)
frac = rt.to_device(pos @ np.linalg.inv(cell), torch.float64)
box = rt.to_device(cell.T.copy(), torch.float64)
dp = rt.DeviceParams(syn.gmnn_params(seed=1234), {"pair_block": variant})
system = rt.PreparedMinimageSystem(dp, Z, int(n * 96) + 4096)
system.build(frac, box, 6.0)
system.check()
for _ in range(steps):
    e, f = system.compute(frac, box, None, "minimage")
torch.cuda.synchronize()
print("E", float(e.cpu()[0]))
