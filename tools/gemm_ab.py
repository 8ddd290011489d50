"""A/B of the readout GEMM variants (GPU box): eight-warp epilogue (shipped) against the sixteen-warp one (apax_i8_set_debug flag 16)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from apax_b200 import runtime as rt, synthetic as syn

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 333334
n_out = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pos, Z, cell = syn.water_box(n_mol, 0)
n = len(Z)
frac = rt.to_device(pos @ np.linalg.inv(cell), torch.float64)
box = rt.to_device(cell.T.copy(), torch.float64)
dp = rt.DeviceParams(syn.gmnn_params(seed=1234, n_out=n_out))
system = rt.PreparedMinimageSystem(dp, Z, int(n * 96) + 4096)
system.build(frac, box, 6.0)
system.check()
ref = None
for flag, name in ((0, "shipped"), (32, "one fewer resident plane (K = 384: 2 resident, 3 x 32 KB stages)"), (64, "two fewer (K = 384: 1 resident, 4 x 40 KB stages, two MMA warps)")):
    rt.lib().apax_i8_set_debug(flag, 0)
    for _ in range(3):
        e, f = system.compute(frac, box, None, "minimage")
    torch.cuda.synchronize()
    rec = rt.profile_kernels(lambda: [system.compute(frac, box, None, "minimage") for _ in range(5)])
    tot = {}
    for k, ms in rec:
        tot[k] = tot.get(k, 0.0) + ms / 5
    gem = {k: round(v, 3) for k, v in tot.items() if "i8_gemm" in k}
    e, f = e.cpu().numpy().copy(), f.cpu().numpy().copy()
    cmp_ = "" if ref is None else f" | dE/E {abs(e[0] - ref[0][0]) / abs(ref[0][0]):.2e} max|dF| {np.abs(f - ref[1]).max():.2e}"
    if ref is None:
        ref = (e, f)
    print(f"{name}: step {sum(tot.values()):.3f} ms, GEMMs {sum(gem.values()):.3f} ms {gem} status {rt.lib().apax_gm_params_status(dp.handle)}{cmp_}", flush=True)
rt.lib().apax_i8_set_debug(0, 0)
