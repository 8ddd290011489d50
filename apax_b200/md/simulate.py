"""Host mirror of the energy-function factory of apax/md/simulate.py (create_energy_fn, :41-74): the
callable JaxMD integrators differentiate.  Here one call returns the energy AND caches the forces of
the same positions (one fused evaluation), which is what `quantity.canonicalize_force`
(apax/utils/jax_md_reduced/quantity.py:68-94) would obtain with jax.grad; INTEGRATION.md shows the
jax.custom_vjp wrapper that exposes exactly this pair to JAX."""
from __future__ import annotations

import torch

from ..nn.models import EnergyModel, canonicalize_neighbors, device_params, displacement_mode
from .. import runtime as rt


def create_energy_fn(model: EnergyModel, params, numbers, n_models: int = 1, shallow: bool = False):
    if n_models > 1 and not shallow:
        raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "create_energy_fn", "full (vmapped-parameter) ensembles")
    # the integrators differentiate the MEAN energy of a shallow ensemble (simulate.py:55-58): one backward pass
    dp = device_params(params, {**model.statics(), "mean_forces": True})
    mode = displacement_mode(model.init_box, model.inference_disp_fn)
    Z = rt.to_device(numbers, torch.int32)
    state = {}

    def energy_fn(R, neighbor, box, offsets=None, **unused):
        e, f = rt.energy_forces(dp, R, Z, canonicalize_neighbors(neighbor), box, offsets, mode)
        state["forces"] = f.mean(dim=0)          # mean over heads == grad of the mean energy (simulate.py:55-58)
        return e.mean()

    def force_fn(R, neighbor, box, offsets=None, **unused):
        energy_fn(R, neighbor, box, offsets)
        return state["forces"]

    energy_fn.force_fn = force_fn
    return energy_fn


def run_nvt(params, atoms, model_config=None, *, dt: float, kT: float, n_steps: int, dr_threshold: float = 0.5,
            rebuild_every: int | None = None, masses=None, momentum=None, seed: int = 0, chain_length: int = 5,
            chain_steps: int = 2, sy_steps: int = 3, tau=None, on_step=None, strict_skin: bool = True):
    """The MD driver of apax/md/simulate.py (md_setup :440-561 + run_sim :218-437) for the NVT Nose-Hoover
    ensemble, with everything inside the step loop on the GPU: fractional positions, min-image list with skin
    (rebuilt on the reference's skin criterion, partition.py:771-779, or every `rebuild_every` steps as
    BASELINE config 3 words it), GMNN energy+forces, integrator.  Units are the caller's (ASE: eV, A, amu =>
    time in A sqrt(amu/eV)).  Returns the final NVTNoseHooverState and a dict of counters.

    With a fixed cadence the list is not consulted between rebuilds, so after every chunk the reference's skin criterion
    (any atom moved more than dr_threshold / 2 since the list was built, partition.py:771-779) is evaluated on the device;
    chunks that violated it are counted in counters["skin_violations"] and, with `strict_skin`, raise at the end of the
    run -- neighbours may have been missed and the trajectory is then not the reference's."""
    import numpy as np

    from ..nn.builder import GMNNBuilder
    from ..utils.jax_md_reduced import partition
    from ..utils.jax_md_reduced.simulate import nvt_nose_hoover

    cell = np.asarray(getattr(atoms.cell, "array", atoms.cell), dtype=np.float64)
    box = cell.T.copy()                                                   # sim_utils.py:18-38
    frac = np.asarray(atoms.positions, dtype=np.float64) @ np.linalg.inv(cell)
    frac = frac - np.floor(frac)
    builder = GMNNBuilder(model_config)
    model = builder.build_energy_model(init_box=box, inference_disp_fn=object())
    r_max = float(builder.config["basis"]["r_max"])
    neighbor_fn = partition.neighbor_list(None, box, r_max, dr_threshold, fractional_coordinates=True,
                                          disable_cell_list=True, format=partition.Sparse)
    energy_fn = create_energy_fn(model, params, atoms.numbers, n_models=1)
    box_d = rt.to_device(box, torch.float64)
    nbrs = neighbor_fn.allocate(frac, box=box_d)
    counters = {"rebuilds": 0, "steps": 0}
    dp = device_params(params, model.statics())
    Z = rt.to_device(atoms.numbers, torch.int32)
    system = {"prepared": rt.PreparedSystem(dp, Z, nbrs.idx), "nbrs": nbrs}

    def force_fn(R, box=None, **kw):
        _, f = system["prepared"].compute(R, box_d, None, "minimage")
        return f[0]

    init_fn, apply_fn = nvt_nose_hoover(force_fn, None, dt, kT, chain_length, chain_steps, sy_steps, tau)
    gen = torch.Generator(device="cuda").manual_seed(seed)
    state = init_fn(gen, frac, mass=np.ones(len(atoms.numbers)) if masses is None else masses, momentum=momentum)

    def maybe_rebuild(step):
        if rebuild_every is not None:
            need = step % rebuild_every == 0 and step > 0
            new = neighbor_fn.allocate(state.position, box=box_d) if need else system["nbrs"]
        else:
            new = system["nbrs"].update(state.position, box=box_d)
            need = new is not system["nbrs"]
            if need and new.did_buffer_overflow:                           # simulate.py:388-394
                new = neighbor_fn.allocate(state.position, box=box_d)
        if need:
            system["nbrs"] = new
            system["prepared"] = rt.PreparedSystem(dp, Z, new.idx, workspace=system["prepared"].ws)
            counters["rebuilds"] += 1

    orig_force = force_fn

    def stepping_force(R, box=None, **kw):
        maybe_rebuild(counters["steps"])
        return orig_force(R)

    init_fn2, apply_fn = nvt_nose_hoover(stepping_force, None, dt, kT, chain_length, chain_steps, sy_steps, tau)
    if rebuild_every is not None and on_step is None and dp.n_out == 1:
        # fixed rebuild cadence: all steps between two list builds go down in ONE library call (apax_md_nvt_run), as the
        # reference runs them inside one jitted fori_loop (apax/md/simulate.py:313-354); the list is built AND prepared on the
        # device in one asynchronous call (apax_gm_prepare_minimage) -- positions never leave the GPU, no COO round trip
        import ctypes as C

        n_atoms = int(state.position.shape[0])
        occupancy = int((nbrs.idx[1] < n_atoms).sum().item())
        ps = rt.PreparedMinimageSystem(dp, Z, int(occupancy * 1.25) + 4096)
        ps.build(state.position, box_d, r_max + dr_threshold)
        ref_pos = state.position.clone()
        e_dev = torch.zeros(1, dtype=torch.float64, device=state.position.device)
        violations = torch.zeros(1, dtype=torch.int32, device=state.position.device)
        overflowed = torch.zeros(1, dtype=torch.int32, device=state.position.device)
        done = 0
        while done < n_steps:
            if done > 0 and done % rebuild_every == 0:
                ps.build(state.position, box_d, r_max + dr_threshold)
                overflowed += ps.overflow.to(torch.int32)
                ref_pos.copy_(state.position)
                counters["rebuilds"] += 1
            k = min(rebuild_every - done % rebuild_every, n_steps - done)
            state.force = state.force.contiguous()
            rt.check(rt.lib().apax_md_nvt_run(rt._stream_ptr(), C.byref(apply_fn.cfg), dp.handle, rt._ptr(state.position),
                                              rt._ptr(state.momentum), rt._ptr(state.force), rt._ptr(state.mass), rt._ptr(ps.Z),
                                              rt._ptr(box_d), ps.n_atoms, ps.n_pairs, rt._ptr(state.chain), rt._ptr(e_dev), ps.wp,
                                              ps.wb, int(k)), "apax_md_nvt_run")
            done += k
            # skin criterion against the positions the current list was built from (no host synchronisation here)
            violations += rt.nl_needs_rebuild(state.position, ref_pos, box_d, dr_threshold)
        counters["steps"] = n_steps
        counters["skin_violations"] = int(violations.item())
        if int(overflowed.item()):
            raise rt._lib.ApaxB200Error(rt._lib.ERR_OVERFLOW, "run_nvt", "neighbour list capacity exceeded during the run")
        if strict_skin and counters["skin_violations"]:
            raise rt._lib.ApaxB200Error(rt._lib.ERR_OVERFLOW, "run_nvt",
                                        f"{counters['skin_violations']} chunk(s) of {rebuild_every} steps moved an atom further than "
                                        f"dr_threshold/2 = {dr_threshold / 2}: lower rebuild_every or use the skin criterion")
        return state, counters
    for _ in range(n_steps):
        counters["steps"] += 1
        state = apply_fn(state, box=box_d)
        if on_step is not None:
            on_step(counters["steps"], state)
    return state, counters
