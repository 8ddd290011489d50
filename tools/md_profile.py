"""Per-kernel device times of one E+F evaluation at the MD size (12,288 atoms, skin list), CUDA events per launch."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from apax_b200 import runtime as rt, synthetic as syn
pos, Z, cell = syn.water_12288(0)
params = syn.gmnn_params(seed=1234)
dp = rt.DeviceParams(params)
box = rt.to_device(cell.T.copy(), torch.float64)
frac = rt.to_device(pos @ np.linalg.inv(cell), torch.float64)
idx, n_found, ov = rt.nl_build_minimage(frac, box, 6.5, int(len(Z) * 130))
ps = rt.PreparedSystem(dp, Z, idx[:, : int(n_found.item())].contiguous())
for _ in range(5): ps.compute(frac, box, None, "minimage")
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(50): ps.compute(frac, box, None, "minimage")
ev1.record(); torch.cuda.synchronize()
print(f"E+F at {len(Z)} atoms, {int(n_found.item())} pairs: {ev0.elapsed_time(ev1)/50:.4f} ms per evaluation")
rec = rt.profile_kernels(lambda: [ps.compute(frac, box, None, "minimage") for _ in range(5)])
agg = {}
for n, ms in rec: agg[n] = agg.get(n, 0.0) + ms / 5
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]): print(f"{v*1e3:8.1f} us  {k}")
print(f"{sum(agg.values())*1e3:8.1f} us  total (event-bracketed, ~3-5 us floor per launch)")
