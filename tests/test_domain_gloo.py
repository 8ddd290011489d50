"""Slab domain decomposition + halo exchange (apax_b200/md/domain.py) on CPU with gloo, world_size 2 and 3:
the same bookkeeping and collectives as the GPU path, with the oracle injected as the local compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apax_b200 import synthetic as syn
from apax_b200.md import domain
from oracle import gmnn_oracle as go
from oracle import neighbor_oracle as no


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_partial(params, Z_local, idx, box):
    """local_compute for domain.evaluate: energy of the centrals only, forces on all local atoms."""
    cfg = go.GMNNConfig()

    def fn(R_local, n_central):
        R = R_local.clone().requires_grad_(True)
        _, aux = go.energy_model(params, R, Z_local, idx, box, torch.zeros((idx.shape[1], 3), dtype=torch.float64),
                                 "minimage", cfg, return_aux=True)
        e = aux["E_i"][:n_central].to(torch.float64).sum()
        (g,) = torch.autograd.grad(e, R)
        return e.detach().reshape(1), -g

    return fn


def _worker(rank, world, port, n_mol, box_scale, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    pos, Z, cell = syn.water_box(n_mol, 5, box_length=box_scale)
    frac = torch.as_tensor(pos @ np.linalg.inv(cell))
    box = torch.as_tensor(cell.T.copy())
    params = syn.gmnn_params(seed=1234)
    dec = domain.decompose(frac, box, 6.0, rank, world)
    ids = dec.local_ids
    R_local = frac[ids].clone()
    R_local[dec.n_owned:] = 777.0  # ghosts must be filled by the forward halo, not by construction
    Z_local = torch.as_tensor(Z.astype(np.int64))[ids]
    true_local = frac[ids]
    pairs = no.jaxmd_sparse_pairs(true_local.numpy(), box.numpy(), 6.0)
    pairs = pairs[:, pairs[1] < dec.n_owned]
    halo = domain.HaloExchanger(dec)
    e, f_owned = domain.evaluate(dec, halo, R_local, _oracle_partial(params, Z_local, torch.as_tensor(pairs.astype(np.int64)), box))
    assert torch.equal(R_local, true_local)  # forward halo delivered exactly the owners' positions
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), owned=dec.owned.numpy(), f=f_owned.numpy(), e=e.numpy(),
             n_ghost=dec.n_local - dec.n_owned)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_mol,L", [(2, 64, None), (3, 150, 18.6)])
def test_slab_decomposition_matches_single_domain(tmp_path, world, n_mol, L):
    mp.spawn(_worker, args=(world, _free_port(), n_mol, L, str(tmp_path)), nprocs=world, join=True)
    pos, Z, cell = syn.water_box(n_mol, 5, box_length=L)
    frac = pos @ np.linalg.inv(cell)
    pairs = no.jaxmd_sparse_pairs(frac, cell.T, 6.0)
    ref = go.energy_forces(syn.gmnn_params(seed=1234), frac, Z, pairs, cell.T, None, "minimage")
    seen = np.zeros(len(Z), dtype=bool)
    F = np.zeros_like(ref["forces"])
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert not seen[d["owned"]].any()
        seen[d["owned"]] = True
        F[d["owned"]] = d["f"]
        assert abs(d["e"][0] - ref["energy"]) <= 1e-6 * abs(ref["energy"])  # all-reduced: every rank holds the total
        assert d["n_ghost"] > 0
    assert seen.all()                                                        # every atom owned exactly once
    assert np.all(np.abs(F - ref["forces"]) <= 1e-5 * np.abs(ref["forces"]) + 2e-6)


def test_decompose_is_consistent_across_ranks():
    pos, Z, cell = syn.water_box(216, 1, box_length=26.0)
    frac = torch.as_tensor(pos @ np.linalg.inv(cell))
    box = torch.as_tensor(cell.T.copy())
    world = 4
    decs = [domain.decompose(frac, box, 6.0, r, world) for r in range(world)]
    for a in range(world):
        for b in range(world):
            assert decs[a].send_counts[b] == decs[b].ghost_owner_counts[a]     # what a sends to b is what b expects from a
    owned = torch.cat([d.owned for d in decs])
    assert torch.equal(torch.sort(owned).values, torch.arange(len(Z)))
    with pytest.raises(ValueError):
        domain.decompose(frac, box, 6.0, 0, 5)                                # slabs thinner than the cutoff
