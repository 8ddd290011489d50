"""Worker of tests/test_gpu_multi.py::test_slab_nvt_matches_single_gpu_trajectory (launched with torch.distributed.run):
every rank runs the domain-decomposed NVT loop; rank 0 also runs the single-GPU loop (run_nvt) from the same initial state and
prints one JSON line with the deviations."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from apax_b200 import synthetic as syn  # noqa: E402
from apax_b200.md.ase_calc import Atoms  # noqa: E402
from apax_b200.md.domain_md import SlabNVT  # noqa: E402
from apax_b200.md.simulate import run_nvt  # noqa: E402


def main():
    n_mol, n_steps, cadence = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    cadence = cadence if cadence > 0 else None
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    json_fd = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pos, Z, cell = syn.water_box(n_mol, 3)
    n = len(Z)
    params = syn.gmnn_params(seed=1234)
    masses = np.where(Z == 8, 15.999, 1.008)
    fs = 0.09822694788464063
    dt, kT = 0.5 * fs, 600.0 * 8.617333262e-5            # warm enough for atoms to cross slab faces within the run
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=(n, 3)) * np.sqrt(masses * kT)[:, None]
    p0 -= p0.mean(0)
    atoms = Atoms(positions=pos, numbers=Z, cell=cell)
    md = SlabNVT(params, atoms, None, dt=dt, kT=kT, dr_threshold=0.5, rebuild_every=cadence, masses=masses, momentum=p0.copy())
    md.run(n_steps)
    Rg, Pg, Fg, e = md.state()
    R_multi, P_multi, e_multi = Rg.cpu().numpy(), Pg.cpu().numpy(), float(e.cpu()[0])
    counters = dict(md.counters)
    md.close()
    out = None
    if rank == 0:
        state, c1 = run_nvt(params, atoms, None, dt=dt, kT=kT, n_steps=n_steps, dr_threshold=0.5, rebuild_every=cadence, masses=masses,
                            momentum=p0.copy(), strict_skin=False)
        R1, P1 = state.position.cpu().numpy(), state.momentum.cpu().numpy()
        d = R_multi - R1
        d -= np.round(d)
        dev_A = np.abs(d @ cell).max()
        moved = np.abs(((R1 - (pos @ np.linalg.inv(cell)) % 1.0 + 0.5) % 1.0 - 0.5) @ cell).max()
        out = {"atoms": n, "steps": n_steps, "max_position_deviation_A": float(dev_A), "max_momentum_deviation": float(np.abs(P_multi - P1).max()),
               "momentum_scale": float(np.abs(P1).max()), "max_displacement_A": float(moved), "counters": counters, "single_rebuilds": c1["rebuilds"],
               "energy": e_multi}
    dist.barrier()
    if rank == 0:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
