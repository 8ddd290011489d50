"""CPU ORACLE (test infrastructure, NOT product code) for the NVT Nose-Hoover-chain integrator that
apax's MD driver runs around the GMNN energy function (apax/md/simulate.py:100-119, 313-354):
jax_md's `nvt_nose_hoover` as vendored in apax/utils/jax_md_reduced/simulate.py.

NumPy restatement with the reference's dtypes: positions / momenta / chain state fp64, time steps and
chain masses rounded to fp32 as the `f32(...)` casts do (simulate.py:219-220, 432, 480-482, 618-622).
The force function is injected.  Parity unpinned against jax itself (not installable here); the
arithmetic follows the source line by line, cited below.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np

# simulate.py:308-327
SUZUKI_YOSHIDA_WEIGHTS = {
    1: [1],
    3: [0.828981543588751, -0.657963087177502, 0.828981543588751],
    5: [0.2967324292201065, 0.2967324292201065, -0.186929716880426, 0.2967324292201065, 0.2967324292201065],
    7: [0.784513610477560, 0.235573213359357, -1.17767998417887, 1.31518632068391, -1.17767998417887,
        0.235573213359357, 0.784513610477560],
}


def f32(x) -> float:
    """Value of jnp.float32(x) as a Python float (exact widening)."""
    return float(np.float32(x))


@dataclass
class Chain:
    xi: np.ndarray      # fp64 [chain_length]
    p_xi: np.ndarray    # fp64
    Q: np.ndarray       # fp32 values held in fp64
    tau: float
    KE: float
    dof: int


def chain_masses(kT: float, tau: float, chain_length: int, dof: int) -> np.ndarray:
    """init_fn / update_chain_mass_fn (simulate.py:428-434, 489-495): Q = kT tau^2 ones(f32); Q[0] *= DOF."""
    Q = (np.float32(kT) * np.float32(tau) ** np.float32(2)) * np.ones(chain_length, dtype=np.float32)
    Q[0] = Q[0] * np.float32(dof)
    return Q.astype(np.float64)


def chain_substep(delta: float, chain: Chain, kT: float):
    """substep_fn (simulate.py:436-473).  Returns the momentum scale factor of this substep."""
    xi, p_xi, Q, KE, dof = chain.xi.copy(), chain.p_xi.copy(), chain.Q, chain.KE, chain.dof
    delta_2 = f32(delta / 2.0)
    delta_4 = f32(delta_2 / 2.0)
    delta_8 = f32(delta_4 / 2.0)
    M = len(xi) - 1
    G = p_xi[M - 1] ** 2 / Q[M - 1] - kT
    p_xi[M] = p_xi[M] + delta_4 * G
    p_old = p_xi.copy()  # the scan reads the pre-update p_xi (closure), carries only the newest value (:448-456)
    carry = p_xi[M]
    for m in range(M - 1, 0, -1):
        G = p_old[m - 1] ** 2 / Q[m - 1] - kT
        scale = np.exp(-delta_8 * carry / Q[m + 1])
        carry = scale * (scale * p_old[m] + delta_4 * G)
        p_xi[m] = carry
    G = 2.0 * KE - dof * kT
    scale = np.exp(-delta_8 * p_xi[1] / Q[1])
    p_xi[0] = scale * (scale * p_xi[0] + delta_4 * G)
    scale = np.exp(-delta_2 * p_xi[0] / Q[0])
    KE = KE * scale ** 2
    p_scale = scale
    xi = xi + delta_2 * p_xi / Q
    G = 2.0 * KE - dof * kT
    p_in = p_xi.copy()  # forward scan also closes over the p_xi present before it (:463-470)
    for m in range(M):
        scale = np.exp(-delta_8 * p_in[m + 1] / Q[m + 1])
        upd = scale * (scale * p_in[m] + delta_4 * G)
        G = upd ** 2 / Q[m] - kT
        p_xi[m] = upd
    p_xi[M] = p_xi[M] + delta_4 * G
    chain.xi, chain.p_xi, chain.KE = xi, p_xi, KE
    return p_scale


def chain_half_step(dt: float, chain: Chain, kT: float, chain_steps: int, sy_steps: int) -> float:
    """half_step_chain_fn (simulate.py:475-487): total factor by which the momenta are multiplied."""
    total = 1.0
    if chain_steps == 1 and sy_steps == 1:
        return chain_substep(dt, chain, kT)
    delta = f32(dt / chain_steps)
    ws = SUZUKI_YOSHIDA_WEIGHTS[sy_steps]
    for i in range(chain_steps * sy_steps):
        total *= chain_substep(f32(delta * ws[i % sy_steps]), chain, kT)
    return total


def kinetic_energy(P: np.ndarray, mass: np.ndarray) -> float:
    """quantity.kinetic_energy: 0.5 * sum(p^2 / m)."""
    return float(0.5 * np.sum(P ** 2 / mass[:, None]))


@dataclass
class NVTState:
    position: np.ndarray   # fractional, wrapped [N,3] fp64
    momentum: np.ndarray   # [N,3] fp64
    force: np.ndarray      # [N,3] fp64 (Cartesian)
    mass: np.ndarray       # [N] fp64
    chain: Chain


def nvt_init(R_frac, momentum, mass, force_fn: Callable, kT: float, dt: float, chain_length=5, tau=None) -> NVTState:
    """init_fn (simulate.py:596-609) with explicitly given momenta (JAX's PRNG cannot be reproduced)."""
    dt = f32(dt)
    tau = f32(dt * 100) if tau is None else f32(tau)
    n = R_frac.shape[0]
    dof = 3 * n                                             # quantity.count_dof
    KE = kinetic_energy(momentum, mass)
    chain = Chain(np.zeros(chain_length), np.zeros(chain_length), chain_masses(kT, tau, chain_length, dof), tau, KE, dof)
    return NVTState(np.array(R_frac, dtype=np.float64), np.array(momentum, dtype=np.float64), force_fn(R_frac),
                    np.array(mass, dtype=np.float64), chain)


def nvt_step(state: NVTState, force_fn: Callable, box: np.ndarray, kT: float, dt: float, chain_steps=2, sy_steps=3):
    """apply_fn (simulate.py:611-637) = chain half step, velocity Verlet (:216-231) with the fractional,
    wrapped shift of periodic_general (space.py:284-309), chain half step."""
    dt = f32(dt)
    dt_2 = f32(dt / 2)
    ch = state.chain
    ch.Q = chain_masses(kT, ch.tau, len(ch.xi), ch.dof)
    s = chain_half_step(dt, ch, kT, chain_steps, sy_steps)
    P = state.momentum * s
    P = P + dt_2 * state.force                                         # momentum_step (:156-162)
    dR = dt * P / state.mass[:, None]                                   # position_step (:165-177)
    inv_box = np.linalg.inv(box)
    R = np.mod(state.position + dR @ inv_box.T, 1.0)                    # transform(inv_box, dR); periodic_shift(1.0, R, dR)
    F = force_fn(R)
    P = P + dt_2 * F
    ch.KE = kinetic_energy(P, state.mass)
    s = chain_half_step(dt, ch, kT, chain_steps, sy_steps)
    P = P * s
    state.position, state.momentum, state.force = R, P, F
    return state


def nvt_invariant(state: NVTState, potential_energy: float, kT: float) -> float:
    """simulate.nvt_nose_hoover_invariant: conserved quantity of the extended system."""
    ch = state.chain
    H = potential_energy + kinetic_energy(state.momentum, state.mass)
    H += ch.p_xi[0] ** 2 / (2 * ch.Q[0]) + ch.dof * kT * ch.xi[0]
    for xi, p, Q in zip(ch.xi[1:], ch.p_xi[1:], ch.Q[1:]):
        H += p ** 2 / (2 * Q) + kT * xi
    return float(H)
