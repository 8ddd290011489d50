"""Host mirror of the NVT integrator apax uses (apax/utils/jax_md_reduced/simulate.py:536-639):
`nvt_nose_hoover(energy_or_force_fn, shift_fn, dt, kT, chain_length, chain_steps, sy_steps, tau)` returns
`(init_fn, apply_fn)` like the reference.  The state lives on the GPU; apply_fn is two C-ABI calls around
the force evaluation (apax_md_nvt_pre_force / apax_md_nvt_post_force) -- no host arithmetic."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from ... import runtime as rt
from ..._lib import MdNvtConfig, check, lib


@dataclass
class NVTNoseHooverState:
    """simulate.py:510-534; `chain` is the opaque device buffer of the Nose-Hoover chain."""
    position: torch.Tensor   # fractional [N,3] f64
    momentum: torch.Tensor   # [N,3] f64
    force: torch.Tensor      # [N,3] f64
    mass: torch.Tensor       # [N] f64
    chain: torch.Tensor

    @property
    def velocity(self):
        return self.momentum / self.mass[:, None]

    def chain_numbers(self) -> dict:
        """Host copy of the chain scalars (xi, p_xi, Q, KE) -- for tests and logging only."""
        raw = self.chain.cpu().numpy().view(np.float64)
        return {"xi": raw[0:16].copy(), "p_xi": raw[16:32].copy(), "Q": raw[32:48].copy(), "KE": float(raw[48])}


def nvt_nose_hoover(energy_or_force_fn: Callable, shift_fn=None, dt: float = 1e-3, kT: float = 1.0, chain_length: int = 5,
                    chain_steps: int = 2, sy_steps: int = 3, tau: Optional[float] = None, **sim_kwargs):
    """`energy_or_force_fn(R_frac, **kwargs)` must return the Cartesian forces [N,3] f64 on the device (e.g.
    `create_energy_fn(...).force_fn`); shift_fn is implied (fractional, wrapped: space.py:284-309)."""
    dt32 = float(np.float32(dt))
    tau = float(np.float32(dt32 * 100)) if tau is None else float(np.float32(tau))
    cfg = MdNvtConfig(chain_length=chain_length, chain_steps=chain_steps, sy_steps=sy_steps, dt=dt, kT=kT, tau=tau)

    def init_fn(key, R, mass=1.0, momentum=None, **kwargs):
        """key: a torch.Generator or None (JAX's PRNG stream is not reproducible here); momenta are drawn from
        the Maxwell-Boltzmann distribution and centred (simulate.py:135-153) unless given."""
        R = rt.to_device(R, torch.float64)
        n = R.shape[0]
        m = rt.to_device(np.broadcast_to(np.asarray(mass, dtype=np.float64), (n,)).copy(), torch.float64)
        if momentum is None:
            g = key if isinstance(key, torch.Generator) else None
            p = torch.sqrt(m * kT)[:, None] * torch.randn((n, 3), dtype=torch.float64, device=R.device, generator=g)
            p = p - p.mean(dim=0, keepdim=True)
        else:
            p = rt.to_device(momentum, torch.float64).clone()
        F = energy_or_force_fn(R, **kwargs).contiguous()
        chain = torch.zeros(lib().apax_md_nvt_state_bytes(), dtype=torch.uint8, device=R.device)
        check(lib().apax_md_nvt_init(rt._stream_ptr(), C.byref(cfg), rt._ptr(p), rt._ptr(m), n, rt._ptr(chain)), "apax_md_nvt_init")
        return NVTNoseHooverState(R.clone(), p, F, m, chain)

    def apply_fn(state: NVTNoseHooverState, box=None, **kwargs):
        n = state.position.shape[0]
        box_d = rt.to_device(box, torch.float64)
        check(lib().apax_md_nvt_pre_force(rt._stream_ptr(), C.byref(cfg), rt._ptr(state.position), rt._ptr(state.momentum),
                                          rt._ptr(state.force), rt._ptr(state.mass), rt._ptr(box_d), n, rt._ptr(state.chain)),
              "apax_md_nvt_pre_force")
        state.force = energy_or_force_fn(state.position, box=box_d, **kwargs).contiguous()
        check(lib().apax_md_nvt_post_force(rt._stream_ptr(), C.byref(cfg), rt._ptr(state.momentum), rt._ptr(state.force),
                                           rt._ptr(state.mass), n, rt._ptr(state.chain)), "apax_md_nvt_post_force")
        return state

    apply_fn.cfg = cfg        # for apax_md_nvt_run (several steps per call)
    return init_fn, apply_fn
