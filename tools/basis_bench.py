"""E+F step time of the 1,000,002-atom water box for the default basis (7 Gaussians, r_max 6) and for apax's config-file
default (BesselBasis, n_basis 16, r_max 5) on the same fused kernels (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from apax_b200 import runtime as rt, synthetic as syn

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 333334
pos, Z, cell = syn.water_box(n_mol, 0)
n = len(Z)
frac = rt.to_device(pos @ np.linalg.inv(cell), torch.float64)
box = rt.to_device(cell.T.copy(), torch.float64)
for name, cfg, r_max in (("gaussian n_basis 7 r_max 6", {}, 6.0), ("bessel n_basis 16 r_max 5", {"n_basis": 16, "r_min": 0.0, "r_max": 5.0, "basis_kind": 2}, 5.0),
                         ("bessel n_basis 16 r_max 6", {"n_basis": 16, "r_min": 0.0, "r_max": 6.0, "basis_kind": 2}, 6.0)):
    dp = rt.DeviceParams(syn.gmnn_params(seed=1234, n_basis=cfg.get("n_basis", 7)), cfg)
    system = rt.PreparedMinimageSystem(dp, Z, int(n * 96) + 4096)
    system.build(frac, box, r_max)
    nf = system.check()
    for _ in range(3):
        e, f = system.compute(frac, box, None, "minimage")
    torch.cuda.synchronize()
    rec = rt.profile_kernels(lambda: [system.compute(frac, box, None, "minimage") for _ in range(5)])
    tot = {}
    for k, ms in rec:
        tot[k] = tot.get(k, 0.0) + ms / 5
    pf = sum(v for k, v in tot.items() if "k_pair_fwd" in k)
    pb = sum(v for k, v in tot.items() if "k_pair_bwd" in k)
    print(f"{name:28s}: {nf / n:6.1f} pairs/atom  step {sum(tot.values()):7.3f} ms  pair_fwd {pf:.3f}  pair_bwd {pb:.3f}  "
          f"E {float(e.cpu()[0]):.6f} finite {bool(torch.isfinite(f).all())}", flush=True)
    del system
    torch.cuda.empty_cache()
