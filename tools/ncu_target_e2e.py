"""Workload for ncu captures of the end-to-end path (GPU box): apax_ctx_calculate with a list rebuild on a water box.
    ncu --set full --import-source on -k regex:<kernels> -c <n> python tools/ncu_target_e2e.py [n_mol]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from apax_b200 import runtime as rt, synthetic as syn

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 333334
pos, Z, cell = syn.water_box(n_mol, 0)
dp = rt.DeviceParams(syn.gmnn_params(seed=1234))
ctx = rt.HostContext(dp, len(Z), 96 * len(Z))
for _ in range(2):
    e, f = ctx.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.0, rebuild=True)
print("E", e)
