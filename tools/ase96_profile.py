"""Per-kernel device times of one 96-atom ASECalculator.calculate() (BASELINE configs[0]; image-list path).
`python tools/ase96_profile.py`"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from apax_b200 import runtime as rt, synthetic as syn
from apax_b200.md.ase_calc import ASECalculator, Atoms

pos, Z, cell = syn.water_box(32, 0)
calc = ASECalculator(syn.gmnn_params(seed=1234))
atoms = Atoms(positions=pos, numbers=Z, cell=cell)
rng = np.random.default_rng(0)
def once():
    atoms.positions = pos + rng.normal(size=pos.shape) * 1e-3
    calc.calculate(atoms)
for _ in range(5): once()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): once()
torch.cuda.synchronize()
print(f"wall {(time.perf_counter() - t0) / 200 * 1e3:.4f} ms per call")
rec = rt.profile_kernels(lambda: [once() for _ in range(10)])
agg = {}
for name, ms in rec:
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values()) / 10
print(f"kernel sum {tot * 1e3:.1f} us per call in {sum(v[0] for v in agg.values()) / 10:.1f} launches")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"  {ms / 10 * 1e3:8.1f} us  {n / 10:4.1f} x  {name}")
