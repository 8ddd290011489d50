"""A/B of the pair-kernel variants (apax_gm_params_set_option(APAX_OPT_PAIR_BLOCK)) on the 1M-atom box (GPU box): per-kernel
milliseconds from CUDA events, and a bitwise comparison of energies / forces against variant 1."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from apax_b200 import runtime as rt, synthetic as syn

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 333334
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 3, 4]   # 1 shipped, 3 all packed, 4 round-1 scalar
skin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
pos, Z, cell = syn.water_box(n_mol, 0)
n = len(Z)
frac = rt.to_device(pos @ np.linalg.inv(cell), torch.float64)
box = rt.to_device(cell.T.copy(), torch.float64)
params = syn.gmnn_params(seed=1234)
ref = None
for v in variants:
    dp = rt.DeviceParams(params, {"pair_block": v})
    system = rt.PreparedMinimageSystem(dp, Z, int(n * (96 if skin == 0 else 125)) + 4096)
    system.build(frac, box, 6.0 + skin)
    system.check()
    for _ in range(3):
        e, f = system.compute(frac, box, None, "minimage")
    torch.cuda.synchronize()
    rec = rt.profile_kernels(lambda: [system.compute(frac, box, None, "minimage") for _ in range(5)])
    tot = {}
    for k, ms in rec:
        tot[k] = tot.get(k, 0.0) + ms / 5
    e, f = e.cpu().numpy().copy(), f.cpu().numpy().copy()
    same = "ref" if ref is None else ("bitwise equal" if (np.array_equal(e, ref[0]) and np.array_equal(f, ref[1])) else
                                      f"DIFFERENT max|dF| {np.abs(f - ref[1]).max():.3e}")
    if ref is None:
        ref = (e, f)
    print(f"variant {v}: pair_fwd {tot['k_pair_fwd<LPA>']:.3f} ms  pair_bwd {tot['k_pair_bwd<LPA>']:.3f} ms  step {sum(tot.values()):.3f} ms  [{same}]", flush=True)
    del system
    torch.cuda.empty_cache()
