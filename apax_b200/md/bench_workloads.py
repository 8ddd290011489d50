"""bench.py workloads besides the headline (`--workload ef`): the other BASELINE configurations measured through the public
mirrors.  Each `measure_*` returns one result line (rank 0; None elsewhere) in the headline's unit, atom-steps/s
(atoms x evaluations / s); `run_*` is the stand-alone `--workload` form that owns stdout's JSON line and the process group.

  ase96     configs[0]: ASECalculator.calculate on the 96-atom periodic water box (image-list path: L = 9.86 A < 2 r_max),
            host positions in, list built every call, energy+forces back on the host -- microseconds per call.
  md        configs[2]: NVT Nose-Hoover MD of the 12 288-atom periodic water box, neighbour list (cutoff 6.0 + 0.5 skin)
            rebuilt every 10 steps, everything inside the step loop on the GPU (apax_b200/md/simulate.py::run_nvt).
  ensemble  configs[4]: shallow-ensemble GMNN (16 heads) uncertainty inference on independent 128-atom periodic cells
            through ASECalculator.batch_eval (multi-image lists + batched E, sigma_E, F, sigma_F); cells sharded by
            structure over ranks when launched with torchrun (no collective on the data path).
These are parity-test configurations first; the lines document where they stand, the headline remains --workload ef.
CPU baselines come from bench.py through `cpu_baseline_fn` (the oracle is checker code: nothing in this package imports it)."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np


def _emit(json_fd, line):
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def _standalone(measure, args, world, rank, local, *rest) -> int:
    import torch

    torch.cuda.set_device(local)
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(args, world, rank, local, *rest)
    if rank == 0 and line is not None:
        _emit(json_fd, line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------- configs[0]
def measure_ase96(args, world, rank, local, ClockSampler, metric, unit, calls: int = 200, cpu_baseline_fn=None):
    """BASELINE configs[0] on the GPU path: the per-call latency of ASECalculator.calculate (rank 0 only)."""
    if rank != 0:
        return None
    import torch

    from .. import runtime as rt
    from .. import synthetic as syn
    from .ase_calc import ASECalculator, Atoms

    torch.cuda.set_device(local)
    lib = rt.lib()
    pos, Z, cell = syn.water_96()
    params = syn.gmnn_params(seed=1234)
    calc = ASECalculator(params, dr_threshold=0.5)
    atoms = Atoms(pos.copy(), Z, cell)
    rng = np.random.default_rng(0)
    for _ in range(10):
        calc.calculate(atoms)
    c0 = lib.apax_kernel_launch_count()
    t0 = time.perf_counter()
    for _ in range(calls):
        atoms.positions = pos + rng.normal(scale=0.01, size=pos.shape)      # a new geometry every call, as an MD / optimiser driver would
        res = calc.calculate(atoms)
    wall = time.perf_counter() - t0
    launches = int(lib.apax_kernel_launch_count() - c0)
    assert np.isfinite(res["energy"]) and np.isfinite(res["forces"]).all()
    n = len(Z)
    line = {"metric": metric, "value": n * calls / wall, "unit": unit, "n_gpus": 1, "steps": calls, "warmup": 10,
            "ms_per_step": wall / calls * 1e3, "us_per_call": wall / calls * 1e6, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[0]: GMNN default E+F on the 96-atom periodic water box through the ASE calculator "
                                   "mirror (image-list path, L = 9.856 A; list rebuilt on the GPU every call as the reference's vesin "
                                   "path does, apax/md/ase_calc.py:386-387)", "atoms": n, "pairs": int(calc.ctx.n_pairs),
                       "timing": "wall clock per calculate(): host positions in, host energy+forces out; latency-bound "
                                 "(~30 dependent launches on 96 atoms)"},
            "gpu_launches": launches,
            "e2e": {"value": n * calls / wall, "unit": unit, "h2d_bytes_per_step": int(n * 28 + 72), "d2h_bytes_per_step": int(n * 24 + 8),
                    "ms_per_step": wall / calls * 1e3}}
    if cpu_baseline_fn is not None:
        line["cpu_baseline"] = cpu_baseline_fn(unit)
    return line


# ------------------------------------------------------------------------------------------- configs[2]
def measure_md(args, world, rank, local, ClockSampler, metric, unit, steps=None, cpu_baseline_fn=None):
    if rank != 0:           # single GPU by specification (replicas only)
        return None
    import torch

    from .. import runtime as rt
    from .. import synthetic as syn
    from .ase_calc import Atoms
    from .simulate import run_nvt

    torch.cuda.set_device(local)
    lib = rt.lib()
    steps = steps or args.steps
    pos, Z, cell = syn.water_12288(0)
    n = len(Z)
    params = syn.gmnn_params(seed=1234)
    masses = np.where(Z == 8, 15.999, 1.008)
    fs = 0.09822694788464063                    # ase.units.fs
    dt, kT = 0.25 * fs, 300.0 * 8.617333262e-5
    atoms = Atoms(positions=pos, numbers=Z, cell=cell)
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=(n, 3)) * np.sqrt(masses * kT)[:, None]
    p0 -= p0.mean(0)

    def go(k):
        return run_nvt(params, atoms, None, dt=dt, kT=kT, n_steps=k, dr_threshold=0.5, rebuild_every=10, masses=masses,
                       momentum=p0.copy(), strict_skin=False)

    go(max(3, args.warmup))
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    c0 = lib.apax_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record()
    state, counters = go(steps)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = ev0.elapsed_time(ev1)
    launches = int(lib.apax_kernel_launch_count() - c0)
    held = sampler.hold(lambda: go(steps), ms, seconds=0.7)          # whole runs of the same length, untimed
    clocks = sampler.stop(window_s=wall + held)
    R = state.position.cpu().numpy()
    assert np.isfinite(R).all() and np.isfinite(state.momentum.cpu().numpy()).all()
    line = {"metric": metric, "value": n * steps / (ms * 1e-3), "unit": unit, "n_gpus": 1, "steps": steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: NVT Nose-Hoover MD with GMNN, 12,288-atom periodic water box, list rebuilt "
                                   "every 10 steps (cutoff 6.0 + 0.5 skin)", "atoms": n, "dt_fs": 0.25, "T_K": 300.0,
                       "rebuilds": counters["rebuilds"], "skin_violations": counters.get("skin_violations"),
                       "md_steps_per_s": steps / (ms * 1e-3),
                       "setup": "run_nvt's own setup (list allocation, chain initialisation) is inside the timed region",
                       "l2": "per-step working set ~0.4 GB exceeds the 126 MB L2"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": n * steps / wall, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "ms_per_step": wall / steps * 1e3,
                    "includes": "wall clock of run_nvt (host loop + device work); positions stay on the device between steps, "
                                "as in the reference's MD loop (apax/md/simulate.py:313-354)"}}
    if cpu_baseline_fn is not None:
        line["cpu_baseline"] = cpu_baseline_fn(unit)
    return line


# ------------------------------------------------------------------------------------------- configs[4]
def measure_ensemble(args, world, rank, local, ClockSampler, metric, unit, cells=None, reps=None, cpu_baseline_fn=None):
    import torch

    from .. import runtime as rt
    from .. import synthetic as syn
    from .ase_calc import ASECalculator, Atoms

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
    lib = rt.lib()
    n_cells_total = cells or args.cells
    per_rank = n_cells_total // world
    seeds = range(rank * per_rank, (rank + 1) * per_rank)          # sharded by structure: no data-path collective
    atoms_list = []
    for sd in seeds:
        p, z, c = syn.cell_128(sd)
        atoms_list.append(Atoms(positions=p, numbers=z, cell=c))
    n_atoms = sum(len(a.numbers) for a in atoms_list)
    params = syn.gmnn_params(seed=1234, n_out=16)
    calc = ASECalculator(params)
    bs = args.ensemble_batch

    def go():
        return calc.batch_eval(atoms_list, batch_size=bs)

    sampler = ClockSampler(local, enabled=(rank == 0))             # started before the warm-up: nvidia-smi needs ~1 s to tick
    sampler.start()
    go()                                                           # warm-up (one full pass)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    c0 = lib.apax_kernel_launch_count()
    reps = max(1, reps or args.steps)
    t0 = time.perf_counter()
    for _ in range(reps):
        res = go()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if dist:
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    launches = int(lib.apax_kernel_launch_count() - c0)
    held = sampler.hold(go, wall / reps * 1e3, seconds=0.7)         # `wall` is the all-reduced figure: same count on every rank
    clocks = sampler.stop(window_s=wall + held)
    assert len(res) == len(atoms_list) and np.isfinite(res[0]["forces_uncertainty"]).all()
    if rank != 0:
        return None
    value = world * n_atoms * reps / wall
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": reps, "warmup": 1,
            "ms_per_step": wall / reps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: shallow-ensemble GMNN (16 readout heads) E, sigma_E, F, sigma_F on independent "
                                   "128-atom periodic cells (42 H2O + H2, L = 10.85 A, multi-image lists), ASECalculator.batch_eval",
                       "cells": per_rank * world, "atoms": world * n_atoms, "batch_size": bs, "parallelism": f"{world} rank(s), cells sharded",
                       "cells_per_s": world * per_rank * reps / wall,
                       "timing": "wall clock over whole batch_eval passes: host cells in, neighbour lists built on the GPU, results "
                                 "(incl. per-head forces) back on the host"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": int(n_atoms * 28 + per_rank * 76),
                    "d2h_bytes_per_step": int(n_atoms * 24 * 16 + per_rank * 8 * 16), "ms_per_step": wall / reps * 1e3}}
    if cpu_baseline_fn is not None:
        line["cpu_baseline"] = cpu_baseline_fn(unit)
    return line


# ------------------------------------------------------------------------------------------- large-system MD (north_star)
def measure_md_slab(args, world, rank, local, ClockSampler, metric, unit, steps: int = 30, molecules=None):
    """NVT MD of the 1,000,002-atom box spread over `world` GPUs (apax_b200/md/domain_md.py::SlabNVT): slab decomposition,
    halo exchange over NVLink peer memory every step, list rebuilt (with atom migration) every 10 steps."""
    import torch
    import torch.distributed as dist

    from .. import synthetic as syn
    from .ase_calc import Atoms
    from .domain_md import SlabNVT

    torch.cuda.set_device(local)
    pos, Z, cell = syn.water_box(molecules or args.molecules, 0)
    n = len(Z)
    params = syn.gmnn_params(seed=1234)
    masses = np.where(Z == 8, 15.999, 1.008)
    fs = 0.09822694788464063
    dt, kT = 0.25 * fs, 300.0 * 8.617333262e-5
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=(n, 3)) * np.sqrt(masses * kT)[:, None]
    p0 -= p0.mean(0)
    md = SlabNVT(params, Atoms(positions=pos, numbers=Z, cell=cell), None, dt=dt, kT=kT, dr_threshold=0.5, rebuild_every=10,
                 masses=masses, momentum=p0)
    md.run(12)                                                      # warm-up: through the first re-decomposition (step 10), whose
    torch.cuda.synchronize()                                        # first-time allocations do not belong in the timed region
    dist.barrier()
    before = dict(md.counters)
    t0 = time.perf_counter()
    md.run(steps)
    torch.cuda.synchronize()
    dist.barrier()
    wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    Rg, Pg, Fg, e = md.state()
    ok = bool(torch.isfinite(Rg).all() and torch.isfinite(Pg).all())
    counters = {k: v - before.get(k, 0) for k, v in md.counters.items()}     # of the timed region
    md.close()
    if rank != 0:
        return None
    wall = float(wall.item())
    return {"metric": metric, "value": n * steps / wall, "unit": unit, "n_gpus": world, "steps": steps, "warmup": 12,
            "ms_per_step": wall / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "north_star large-system MD: NVT Nose-Hoover MD with GMNN on the 1,000,002-atom periodic water box, "
                                   "slab decomposition over the GPUs, halo exchange over NVLink peer memory every step, list (6.0 + 0.5 A "
                                   "skin) rebuilt with atom migration every 10 steps", "atoms": n, "parallelism": f"{world} slabs along x",
                       "rebuilds": counters["rebuilds"], "migrated_atoms": counters["migrated"], "finite": ok,
                       "timing": "wall clock over the step loop incl. the list rebuilds inside it, max over ranks"},
            "e2e": {"value": n * steps / wall, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "ms_per_step": wall / steps * 1e3}}


def run_md(args, world, rank, local, ClockSampler, metric, unit, cpu_baseline_fn=None) -> int:
    return _standalone(lambda *a: measure_md(*a, cpu_baseline_fn=cpu_baseline_fn), args, world, rank, local, ClockSampler, metric, unit)


def run_ensemble(args, world, rank, local, ClockSampler, metric, unit, cpu_baseline_fn=None) -> int:
    return _standalone(lambda *a: measure_ensemble(*a, cpu_baseline_fn=cpu_baseline_fn), args, world, rank, local, ClockSampler, metric, unit)


def run_ase96(args, world, rank, local, ClockSampler, metric, unit, cpu_baseline_fn=None) -> int:
    return _standalone(lambda *a: measure_ase96(*a, cpu_baseline_fn=cpu_baseline_fn), args, world, rank, local, ClockSampler, metric, unit)
