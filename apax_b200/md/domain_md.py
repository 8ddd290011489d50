"""NVT Nose-Hoover molecular dynamics of ONE periodic box spread over the GPUs of a node (north_star: "Large-system MD:
spatial domain decomposition across the 8 GPUs with NVLink halo exchange").  The reference has no counterpart -- its MD loop
runs one structure on one device (apax/md/simulate.py:313-354) -- so what must be matched is the TRAJECTORY of that loop:
the same integrator (jax_md's nvt_nose_hoover, apax/utils/jax_md_reduced/simulate.py:536-639), the same neighbour-list
semantics (cutoff r_max + skin, rebuilt on the skin criterion partition.py:771-779 or at a fixed cadence), the same forces.

Per step and rank (owned atoms only; all kernels from libapax_b200, nothing computed in Python):
  chain half step (replicated, driven by the GLOBAL kinetic energy)      apax_md_nvt_pre_force
  kick + drift of the owned atoms
  forward halo: new positions of my boundary atoms -> my neighbours' ghost rows   apax_halo_push_rows (NVLink peer stores)
  E+F of my centrals against owned + ghost atoms                         apax_gm_compute_partial
  reverse halo: ghost force contributions back to their owners           apax_halo_pull_add_rows
  kick, local kinetic energy                                             apax_md_nvt_kick_ke
  kinetic energy summed over ranks in rank order (bitwise equal everywhere)   apax_peer_allreduce_sum
  chain half step + momentum rescale                                     apax_md_nvt_chain_finish
At a list rebuild atoms may have left their slab: every rank publishes its owned (id, R, P, F) rows (one all-gather over
NVLink), the decomposition is re-derived from the gathered positions -- a pure function of them, hence identical on every
rank -- ownership and ghosts follow, the peer blocks get their new row layout (allocated once, with head-room), and the
neighbour list is rebuilt on the device (apax_gm_prepare_minimage).  Momenta and forces travel with the atoms, so the
trajectory does not notice the hand-over."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import domain


def gather_owned_rows(dec, group, world: int, R_owned, P_owned, F_owned, Rg, Pg, Fg) -> None:
    """Every rank's owned (id, R, P, F) rows into the replicated global arrays Rg / Pg / Fg [N,3] (one all-gather; rows
    padded to the largest owner).  Pure torch + torch.distributed: the CPU twin of the MD loop (tests, gloo) uses the same code."""
    no = dec.n_owned
    counts = [None] * world
    dist.all_gather_object(counts, no, group=group)
    m = max(counts)
    buf = torch.zeros((m, 10), dtype=torch.float64, device=Rg.device)
    buf[:no, 0] = dec.owned.to(torch.float64)
    buf[:no, 1:4] = R_owned
    buf[:no, 4:7] = P_owned
    buf[:no, 7:10] = F_owned
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    for r, part in enumerate(parts):
        ids = part[: counts[r], 0].to(torch.int64)
        Rg[ids] = part[: counts[r], 1:4]
        Pg[ids] = part[: counts[r], 4:7]
        Fg[ids] = part[: counts[r], 7:10]


def redistribute_from_global(Rg, Pg, Fg, Mg, box, cutoff: float, rank: int, world: int):
    """Ownership and ghosts re-derived from the gathered positions (a pure function of them: identical on every rank) and the
    state that travels with the atoms: (decomposition, local positions [owned, then ghosts], owned momenta / forces / masses)."""
    dec = domain.decompose(Rg, box, cutoff, rank, world)
    return dec, Rg[dec.local_ids].clone(), Pg[dec.owned].contiguous(), Fg[dec.owned].contiguous(), Mg[dec.owned].contiguous()


class SlabNVT:
    def __init__(self, params: dict, atoms, model_config: Optional[dict] = None, *, dt: float, kT: float, group=None,
                 dr_threshold: float = 0.5, rebuild_every: Optional[int] = None, masses=None, momentum=None, chain_length: int = 5,
                 chain_steps: int = 2, sy_steps: int = 3, tau: Optional[float] = None, capacity_factor: float = 1.5,
                 pairs_per_atom: int = 130):
        from .. import runtime as rt
        from .._lib import MdNvtConfig
        from ..nn.builder import GMNNBuilder
        from ..nn.models import device_params

        self.rt, self.lib, self.group = rt, rt.lib(), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        cell = np.asarray(getattr(atoms.cell, "array", atoms.cell), dtype=np.float64)
        box = cell.T.copy()                                                    # sim_utils.py:18-38
        frac = np.asarray(atoms.positions, dtype=np.float64) @ np.linalg.inv(cell)
        frac = frac - np.floor(frac)
        builder = GMNNBuilder(model_config)
        model = builder.build_energy_model(init_box=box, inference_disp_fn=object())
        self.dp = device_params(params, model.statics())
        if self.dp.n_out != 1:
            raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "SlabNVT", "single-head models only")
        self.r_max = float(builder.config["basis"]["r_max"])
        self.skin, self.rebuild_every = float(dr_threshold), rebuild_every
        self.n = n = frac.shape[0]
        self.box = torch.as_tensor(box, dtype=torch.float64, device=dev)
        # replicated global state (valid at rebuild points): positions, momenta, forces by global atom id
        self.Rg = torch.as_tensor(frac, dtype=torch.float64, device=dev)
        self.Pg = torch.as_tensor(np.zeros((n, 3)) if momentum is None else np.asarray(momentum, dtype=np.float64), device=dev).contiguous()
        self.Fg = torch.zeros((n, 3), dtype=torch.float64, device=dev)
        self.Mg = torch.as_tensor(np.ones(n) if masses is None else np.asarray(masses, dtype=np.float64), device=dev)
        self.Zg = torch.as_tensor(np.asarray(atoms.numbers), dtype=torch.int32, device=dev)
        self.cfg = MdNvtConfig(chain_length=chain_length, chain_steps=chain_steps, sy_steps=sy_steps, dt=float(dt), kT=float(kT),
                               tau=float(tau if tau is not None else 100.0 * dt))
        self.chain = torch.zeros(self.lib.apax_md_nvt_state_bytes(), dtype=torch.uint8, device=dev)
        self.ke_local = torch.zeros(1, dtype=torch.float64, device=dev)
        self.ke_global = torch.zeros(1, dtype=torch.float64, device=dev)
        self.flag_d = torch.zeros(1, dtype=torch.float64, device=dev)
        self.flag_g = torch.zeros(1, dtype=torch.float64, device=dev)
        self.e_local = None
        self.capacity_factor, self.pairs_per_atom = capacity_factor, pairs_per_atom
        self.peer: Optional[domain.PeerHalo] = None
        self.system = None
        self.workspaces = (rt.Workspace(), rt.Workspace())
        self.counters = {"steps": 0, "rebuilds": 0, "migrated": 0}
        self._redistribute(first=True)
        # forces of the initial positions, kinetic energy of the initial momenta
        self._forces()
        s = rt._stream_ptr()
        no = self.dec.n_owned
        rt.check(self.lib.apax_md_nvt_init(s, C.byref(self.cfg), rt._ptr(self.P), rt._ptr(self.mass), max(no, 1), rt._ptr(self.chain)),
                 "apax_md_nvt_init")
        rt.check(self.lib.apax_md_nvt_kick_ke(s, C.byref(self.cfg), rt._ptr(self.P), None, rt._ptr(self.mass), no, 0, rt._ptr(self.chain),
                                              rt._ptr(self.ke_local)), "apax_md_nvt_kick_ke")
        self.peer.allreduce_sum(self.ke_local, self.ke_global)
        rt.check(self.lib.apax_md_nvt_set_global(s, rt._ptr(self.chain), rt._ptr(self.ke_global), self.n), "apax_md_nvt_set_global")

    # ------------------------------------------------------------------------------------------ data movement
    def _redistribute(self, first: bool = False) -> None:
        """Ownership, ghosts, peer layout and neighbour list from the gathered global state."""
        import time

        rt = self.rt
        prof = getattr(self, "profile", None)      # optional {phase: seconds} accumulator (synchronising; measurement only)

        def mark(name, t0):
            if prof is not None:
                torch.cuda.synchronize()
                prof[name] = prof.get(name, 0.0) + time.perf_counter() - t0
            return time.perf_counter()

        t = time.perf_counter()
        dec, R_loc, P_own, F_own, M_own = redistribute_from_global(self.Rg, self.Pg, self.Fg, self.Mg, self.box, self.r_max + self.skin,
                                                                   self.rank, self.world)
        t = mark("decompose", t)
        if first:
            cap_rows = int(dec.n_local * self.capacity_factor) + 1024
            self.peer = domain.PeerHalo(dec, self.group, capacity_rows=cap_rows)
        else:
            self.peer.configure(dec, barrier=False)
            # atoms I own now that I did not own before (ownership kept as a per-atom flag: no set operations on 10^5 ids)
            self.counters["migrated"] += int((~self.owned_flag[dec.owned]).sum().item())
        t = mark("configure", t)
        self.dec = dec
        self.owned_flag = torch.zeros(self.n, dtype=torch.bool, device=self.Rg.device)
        self.owned_flag[dec.owned] = True
        ids = dec.local_ids
        no = dec.n_owned
        self.R = self.peer.R                                   # [n_local, 3] in the peer block: owned rows, then ghost rows
        self.R.copy_(R_loc)
        self.F = self.peer.F                                   # [1, n_local, 3]
        self.F[0, :no].copy_(F_own)
        self.P = P_own
        self.mass = M_own
        self.R_ref = self.R[:no].clone()
        t = mark("copies", t)
        cap = int(dec.n_local * self.pairs_per_atom) + 4096
        self.system = rt.PreparedMinimageSystem(self.dp, self.Zg[ids].contiguous(), cap, n_central=no, forces=self.F,
                                                workspaces=self.workspaces)
        t = mark("system", t)
        self.system.build(self.R, self.box, self.r_max + self.skin)
        self.system.check()
        t = mark("list", t)
        # device-level rendezvous: no peer may push next step's ghost positions into my block before the copy above has run
        # (the host barrier inside configure() orders the host code only)
        self.peer.allreduce_sum(self.flag_d, self.flag_g)
        t = mark("rendezvous", t)
        self.counters["rebuilds"] += 0 if first else 1

    def _gather(self) -> None:
        """Every rank's owned (id, R, P, F) rows into the replicated global arrays."""
        no = self.dec.n_owned
        gather_owned_rows(self.dec, self.group, self.world, self.R[:no], self.P, self.F[0, :no], self.Rg, self.Pg, self.Fg)

    def _forces(self) -> None:
        self.peer.forward()
        e_local, _ = self.system.compute(self.R, self.box, None, "minimage")
        self.energy, _ = self.peer.reverse(e_local)

    # ------------------------------------------------------------------------------------------ the loop
    def _needs_rebuild(self, step: int) -> bool:
        if self.rebuild_every is not None:
            return step > 0 and step % self.rebuild_every == 0
        flag = self.rt.nl_needs_rebuild(self.R[: self.dec.n_owned], self.R_ref, self.box, self.skin)
        self.flag_d.copy_(flag.to(torch.float64))
        self.peer.allreduce_sum(self.flag_d, self.flag_g)        # any rank's violation rebuilds everywhere
        return bool(self.flag_g.item() > 0)

    def run(self, n_steps: int):
        rt, lib = self.rt, self.lib
        cfg = C.byref(self.cfg)
        for _ in range(n_steps):
            step = self.counters["steps"]
            if self._needs_rebuild(step):
                self._gather()
                self._redistribute()
            s = rt._stream_ptr()
            no = self.dec.n_owned
            rt.check(lib.apax_md_nvt_pre_force(s, cfg, rt._ptr(self.R), rt._ptr(self.P), rt._ptr(self.F), rt._ptr(self.mass),
                                               rt._ptr(self.box), no, rt._ptr(self.chain)), "apax_md_nvt_pre_force")
            self._forces()
            rt.check(lib.apax_md_nvt_kick_ke(s, cfg, rt._ptr(self.P), rt._ptr(self.F), rt._ptr(self.mass), no, 1, rt._ptr(self.chain),
                                             rt._ptr(self.ke_local)), "apax_md_nvt_kick_ke")
            self.peer.allreduce_sum(self.ke_local, self.ke_global)
            rt.check(lib.apax_md_nvt_chain_finish(s, cfg, rt._ptr(self.P), no, rt._ptr(self.chain), rt._ptr(self.ke_global), self.n),
                     "apax_md_nvt_chain_finish")
            self.counters["steps"] += 1
        return self

    def state(self):
        """Gathered global (fractional positions, momenta, forces) [N,3] f64 tensors and the potential energy [1]."""
        self._gather()
        torch.cuda.synchronize()
        rt = self.rt
        rt.check(self.lib.apax_halo_status(C.byref(self.peer.fwd)), "apax_halo_status")
        return self.Rg, self.Pg, self.Fg, self.energy

    def close(self):
        if self.peer is not None:
            dist.barrier(group=self.group)
            self.peer.close()
            self.peer = None
