"""One E+F evaluation of the 1M-atom box with apax_i8_set_debug(0, 1): per-launch cycle counters of the i8 GEMM roles."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from apax_b200 import runtime as rt, synthetic as syn

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 333334
pos, Z, cell = syn.water_box(n_mol, seed=0)
rt.lib().apax_i8_set_debug(0, 1)
dp = rt.DeviceParams(syn.gmnn_params(seed=1))
ctx = rt.HostContext(dp, len(Z), 96 * len(Z))
for _ in range(2):
    e, f = ctx.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.0, rebuild=True)
print("E", e)
