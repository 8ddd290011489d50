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
    dp = device_params(params, model.statics())
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
