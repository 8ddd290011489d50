"""Where an MD step at 12,288 atoms goes (GPU box): CUDA-event bracketed kernels over 50 steps of run_nvt (cadence 10)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from apax_b200 import runtime as rt, synthetic as syn
from apax_b200.md.ase_calc import Atoms
from apax_b200.md.simulate import run_nvt

pos, Z, cell = syn.water_12288(0)
n = len(Z)
params = syn.gmnn_params(seed=1234)
masses = np.where(Z == 8, 15.999, 1.008)
fs = 0.09822694788464063
dt, kT = 0.25 * fs, 300.0 * 8.617333262e-5
rng = np.random.default_rng(0)
p0 = rng.normal(size=(n, 3)) * np.sqrt(masses * kT)[:, None]
p0 -= p0.mean(0)
atoms = Atoms(positions=pos, numbers=Z, cell=cell)
go = lambda k: run_nvt(params, atoms, None, dt=dt, kT=kT, n_steps=k, dr_threshold=0.5, rebuild_every=10, masses=masses, momentum=p0.copy(), strict_skin=False)
go(10)
torch.cuda.synchronize()
t0 = time.perf_counter()
go(100)
torch.cuda.synchronize()
print(f"wall {1e3 * (time.perf_counter() - t0) / 100:.4f} ms/step over 100 steps")
steps = 50
rec = rt.profile_kernels(lambda: go(steps), max_records=100000)
agg, cnt = {}, {}
for k, ms in rec:
    agg[k] = agg.get(k, 0.0) + ms / steps
    cnt[k] = cnt.get(k, 0) + 1
print(f"kernel sum {sum(agg.values()):.4f} ms/step in {len(rec) / steps:.1f} launches/step")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:30]:
    print(f"  {v * 1e3:8.1f} us/step  {cnt[k] / steps:5.1f} x  {k}")
