"""Where does a shallow-ensemble batch_eval pass spend its time?  (GPU box)  Kernel sums per pass (CUDA events around every
launch) against the wall clock, with and without the per-head forces travelling to the host."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from apax_b200 import runtime as rt, synthetic as syn
from apax_b200.md.ase_calc import ASECalculator, Atoms

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
atoms = [Atoms(*syn.cell_128(s)) for s in range(n_cells)]
calc = ASECalculator(syn.gmnn_params(seed=1234, n_out=16))
calc.batch_eval(atoms, batch_size=bs)
torch.cuda.synchronize()
t0 = time.perf_counter()
calc.batch_eval(atoms, batch_size=bs)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
rec = rt.profile_kernels(lambda: calc.batch_eval(atoms, batch_size=bs), max_records=200000)
tot = {}
for k, ms in rec:
    tot[k] = tot.get(k, 0.0) + ms
print(f"{n_cells} cells, batch {bs}: wall {wall * 1e3:.1f} ms per pass, kernels {sum(tot.values()):.1f} ms in {len(rec)} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
    print(f"  {k:40s} {v:8.2f} ms")
