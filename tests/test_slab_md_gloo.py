"""CPU twin (gloo, world size 2) of the domain-decomposed NVT loop (apax_b200/md/domain_md.py): the SAME bookkeeping --
`domain.decompose`, the halo exchanger, `gather_owned_rows`, `redistribute_from_global` -- driven by the oracle's force field
and the oracle's restatement of jax_md's nvt_nose_hoover (oracle/md_oracle.py), against the single-process oracle
trajectory.  What it pins down: momenta and forces travel with the atoms when they change owner at a list rebuild, the
replicated chain sees the global kinetic energy, ghosts are re-derived consistently on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apax_b200 import synthetic as syn
from apax_b200.md import domain
from apax_b200.md.domain_md import gather_owned_rows, redistribute_from_global
from oracle import gmnn_oracle as go
from oracle import md_oracle as mo
from oracle import neighbor_oracle as no

R_MAX, SKIN, CADENCE, STEPS = 6.0, 0.5, 3, 9
KT, DT = 0.35, 0.12


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _system():
    pos, Z, cell = syn.water_box(100, 4)                       # L = 14.4 A: two slabs of 7.2 A >= 6.5 A
    masses = np.where(Z == 8, 15.999, 1.008)
    rng = np.random.default_rng(1)
    p0 = rng.normal(size=(len(Z), 3)) * np.sqrt(masses * KT)[:, None]
    p0 -= p0.mean(0)
    frac = pos @ np.linalg.inv(cell)
    frac -= np.floor(frac)
    return frac, Z, cell, masses, p0, syn.gmnn_params(seed=5)


def _forces_central(params, R_local, Z_local, pairs, box, n_central):
    """fp64 oracle: energy of the centrals, forces on all local atoms (what apax_gm_compute_partial returns)."""
    cfg = go.fp64_config()
    R = R_local.clone().requires_grad_(True)
    idx = torch.as_tensor(pairs.astype(np.int64))
    _, aux = go.energy_model(params, R, Z_local, idx, box, torch.zeros((idx.shape[1], 3), dtype=torch.float64), "minimage", cfg,
                             return_aux=True)
    e = aux["E_i"][:n_central].to(torch.float64).sum()
    (g,) = torch.autograd.grad(e, R)
    return -g


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    frac, Z, cell, masses, p0, params = _system()
    n = len(Z)
    box = torch.as_tensor(cell.T.copy())
    Rg, Pg, Fg = torch.as_tensor(frac.copy()), torch.as_tensor(p0.copy()), torch.zeros((n, 3), dtype=torch.float64)
    Mg, Zg = torch.as_tensor(masses), torch.as_tensor(Z.astype(np.int64))
    dt, dt_2 = mo.f32(DT), mo.f32(mo.f32(DT) / 2)
    chain = None
    migrated = 0

    def redistribute():
        dec, R_loc, P, F, M = redistribute_from_global(Rg, Pg, Fg, Mg, box, R_MAX + SKIN, rank, world)
        pairs = no.jaxmd_sparse_pairs(R_loc.numpy(), box.numpy(), R_MAX + SKIN)
        pairs = pairs[:, pairs[1] < dec.n_owned]
        return dec, domain.HaloExchanger(dec), R_loc, P, F, M, Zg[dec.local_ids], pairs

    def forces(dec, halo, R_loc, Z_loc, pairs):
        halo.forward(R_loc)
        f_local = _forces_central(params, R_loc, Z_loc, pairs, box, dec.n_owned)
        return halo.reverse(f_local.clone())

    def global_ke(P, M):
        ke = torch.tensor([0.5 * float(torch.sum(P ** 2 / M[:, None]))], dtype=torch.float64)
        dist.all_reduce(ke)
        return float(ke[0])

    dec, halo, R_loc, P, F, M, Z_loc, pairs = redistribute()
    F = forces(dec, halo, R_loc, Z_loc, pairs).clone()
    tau = mo.f32(dt * 100)
    chain = mo.Chain(np.zeros(5), np.zeros(5), mo.chain_masses(KT, tau, 5, 3 * n), tau, global_ke(P, M), 3 * n)
    inv_box = torch.linalg.inv(box)
    for step in range(STEPS):
        if step > 0 and step % CADENCE == 0:
            gather_owned_rows(dec, None, world, R_loc[: dec.n_owned], P, F, Rg, Pg, Fg)
            old = set(dec.owned.tolist())
            dec, halo, R_loc, P, F, M, Z_loc, pairs = redistribute()
            migrated += len(set(dec.owned.tolist()) - old)
        chain.Q = mo.chain_masses(KT, chain.tau, 5, chain.dof)
        s = mo.chain_half_step(dt, chain, KT, 2, 3)
        P = P * s + dt_2 * F
        dR = dt * P / M[:, None]
        no_ = dec.n_owned
        R_loc[:no_] = torch.remainder(R_loc[:no_] + dR @ inv_box.T, 1.0)
        F = forces(dec, halo, R_loc, Z_loc, pairs).clone()
        P = P + dt_2 * F
        chain.KE = global_ke(P, M)
        s = mo.chain_half_step(dt, chain, KT, 2, 3)
        P = P * s
    gather_owned_rows(dec, None, world, R_loc[: dec.n_owned], P, F, Rg, Pg, Fg)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), R=Rg.numpy(), P=Pg.numpy(), F=Fg.numpy(), migrated=migrated, xi=chain.xi, p_xi=chain.p_xi)
    dist.destroy_process_group()


def test_slab_nvt_bookkeeping_matches_single_process_trajectory(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    frac, Z, cell, masses, p0, params = _system()
    box = cell.T.copy()
    lists = {}

    def force_fn_at(step):
        def fn(R):
            key = (step // CADENCE) if step >= 0 else 0
            if key not in lists:                               # the list is rebuilt from the positions BEFORE the step's drift
                lists[key] = no.jaxmd_sparse_pairs(np.asarray(lists["R_build"]), box, R_MAX + SKIN)
            return go.energy_forces(params, R, Z, lists[key], box, None, "minimage", go.fp64_config())["forces"]
        return fn

    lists["R_build"] = frac.copy()
    state = mo.nvt_init(frac, p0, masses, force_fn_at(0), KT, DT)
    for step in range(STEPS):
        if step > 0 and step % CADENCE == 0:
            lists["R_build"] = state.position.copy()
        state = mo.nvt_step(state, force_fn_at(step), box, KT, DT)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for k in ("R", "P", "F", "xi", "p_xi"):
        assert np.array_equal(got[0][k], got[1][k]), k         # replicated state is bitwise identical on both ranks
    d = got[0]["R"] - state.position
    d -= np.round(d)
    assert np.abs(d @ cell).max() <= 1e-10, np.abs(d @ cell).max()
    assert np.abs(got[0]["P"] - state.momentum).max() <= 1e-10 * np.abs(state.momentum).max()
    assert np.abs(got[0]["F"] - state.force).max() <= 1e-9 * np.abs(state.force).max()
    assert np.allclose(got[0]["xi"], state.chain.xi, rtol=0, atol=1e-12) and np.allclose(got[0]["p_xi"], state.chain.p_xi, rtol=1e-11, atol=1e-12)
    assert int(got[0]["migrated"]) + int(got[1]["migrated"]) > 0   # atoms really changed owner during the run
