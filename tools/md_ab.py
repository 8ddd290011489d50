import time, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
from apax_b200 import synthetic as syn
from apax_b200.md.ase_calc import Atoms
from apax_b200.md.simulate import run_nvt
pos, Z, cell = syn.water_12288(0)
params = syn.gmnn_params(seed=1234)
masses = np.where(Z == 8, 15.999, 1.008)
fs = 0.09822694788464063
dt, kT = 0.25 * fs, 300.0 * 8.617333262e-5
rng = np.random.default_rng(0)
p0 = rng.normal(size=(len(Z), 3)) * np.sqrt(masses * kT)[:, None]; p0 -= p0.mean(0)
atoms = Atoms(positions=pos, numbers=Z, cell=cell)
def go(steps, slow):
    kw = dict(dt=dt, kT=kT, n_steps=steps, dr_threshold=0.5, rebuild_every=10, masses=masses, momentum=p0.copy())
    if slow: kw['on_step'] = lambda i, s: None
    return run_nvt(params, atoms, None, **kw)
for slow in (True, False, True, False):
    go(20, slow); torch.cuda.synchronize()
    t0 = time.perf_counter(); go(200, slow); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print("per-step loop" if slow else "apax_md_nvt_run", f"{t/200*1e3:.3f} ms/step")
