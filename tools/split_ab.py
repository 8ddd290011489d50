"""A/B of the split contraction kernels (APAX_OPT_CONTRACT_SPLIT): E+F time per call with the thread-per-atom kernels
against one-atom-on-four-warps, resident list, at several system sizes.  `python tools/split_ab.py`"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from apax_b200 import runtime as rt, synthetic as syn


out = {}
for n_mol in (32, 342, 1366, 4096, 10923, 21846):
    pos, Z, cell = syn.water_box(n_mol, 0)
    n = len(Z)
    row = {}
    for split in (0, 1 << 22):
        dp = rt.DeviceParams(syn.gmnn_params(seed=1234), {"contract_split": split})
        ctx = rt.HostContext(dp, n, 170 * n)
        ctx.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.5, rebuild=True)
        t = []
        for _ in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); ev0.record()
            for _ in range(20):
                ctx.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.5, rebuild=False)
            ev1.record(); torch.cuda.synchronize()
            t.append(ev0.elapsed_time(ev1) / 20)
        row["split" if split else "thread_per_atom"] = min(t)
        ctx.close()
    out[n] = row
    print(n, row, flush=True)
