"""Seeded synthetic inputs for the GMNN energy+forces hot path.

Everything here is NumPy on the host and deterministic given the seed, so the
oracle, the CUDA path, the tests and bench.py all see identical arrays.  Nothing
in this file touches the GPU or the oracle.
"""
from __future__ import annotations

import numpy as np

# 0.997 g/cm^3 water -> volume per molecule in A^3
WATER_VOL_PER_MOL = 29.915
OH_BOND = 0.9572
HOH_ANGLE_DEG = 104.52


def _random_rotations(rng: np.random.Generator, n: int) -> np.ndarray:
    """n uniformly random rotation matrices [n,3,3] (QR of Gaussian matrices)."""
    a = rng.normal(size=(n, 3, 3))
    q, r = np.linalg.qr(a)
    d = np.sign(np.diagonal(r, axis1=1, axis2=2))
    q = q * d[:, None, :]
    det = np.linalg.det(q)
    q[:, :, 0] *= det[:, None]
    return q


def water_box(n_mol: int, seed: int = 0, extra_h2: bool = False, box_length: float | None = None):
    """Cubic periodic box of rigid water molecules at liquid density.

    Returns (positions [N,3] f64 Cartesian, numbers [N] int32, cell [3,3] f64 with
    lattice vectors as rows, i.e. the ASE convention).  Atom order is O,H,H per
    molecule.  With `extra_h2` one H2 molecule is appended (the 128-atom cells of
    BASELINE config 5: 42 H2O + H2).
    """
    rng = np.random.default_rng(seed)
    n_sites_mol = n_mol + (1 if extra_h2 else 0)
    L = float((n_sites_mol * WATER_VOL_PER_MOL) ** (1.0 / 3.0)) if box_length is None else float(box_length)
    g = int(np.ceil(n_sites_mol ** (1.0 / 3.0) - 1e-9))
    sites = rng.permutation(g ** 3)[:n_sites_mol]
    ix, iy, iz = sites // (g * g), (sites // g) % g, sites % g
    spacing = L / g
    centres = (np.stack([ix, iy, iz], axis=1) + 0.5) * spacing
    centres = centres + rng.uniform(-0.3, 0.3, size=centres.shape)

    half = np.deg2rad(HOH_ANGLE_DEG) / 2.0
    h1 = OH_BOND * np.array([np.sin(half), 0.0, np.cos(half)])
    h2 = OH_BOND * np.array([-np.sin(half), 0.0, np.cos(half)])
    rot = _random_rotations(rng, n_sites_mol)

    pos = np.empty((n_mol, 3, 3))
    pos[:, 0] = centres[:n_mol]
    pos[:, 1] = centres[:n_mol] + rot[:n_mol] @ h1
    pos[:, 2] = centres[:n_mol] + rot[:n_mol] @ h2
    positions = pos.reshape(-1, 3)
    numbers = np.tile(np.array([8, 1, 1], dtype=np.int32), n_mol)
    if extra_h2:
        c = centres[n_mol]
        axis = rot[n_mol] @ np.array([0.0, 0.0, 0.37])
        positions = np.concatenate([positions, (c + axis)[None], (c - axis)[None]], axis=0)
        numbers = np.concatenate([numbers, np.array([1, 1], dtype=np.int32)])
    positions = np.mod(positions, L)
    cell = np.eye(3) * L
    return np.ascontiguousarray(positions), numbers, cell


# named sizes from BASELINE.json / SURVEY.md §8d
def water_96(seed: int = 0):
    return water_box(32, seed)


def water_12288(seed: int = 0):
    return water_box(4096, seed)


def water_1m(seed: int = 0):
    """"1 M" = 1 000 002 atoms (333 334 molecules; 10^6 is not divisible by 3)."""
    return water_box(333_334, seed)


def cell_128(seed: int = 0):
    """42 H2O + one H2 in L = 10.847 A (BASELINE config 5)."""
    return water_box(42, seed, extra_h2=True)


ETHANOL_Z = np.array([6, 6, 8, 1, 1, 1, 1, 1, 1], dtype=np.int32)
ETHANOL_R0 = np.array(
    [
        [0.0072, -0.5687, 0.0000],
        [-1.2854, 0.2499, 0.0000],
        [1.1304, 0.3147, 0.0000],
        [0.0392, -1.1972, 0.8900],
        [0.0392, -1.1972, -0.8900],
        [-1.3175, 0.8784, 0.8900],
        [-1.3175, 0.8784, -0.8900],
        [-2.1422, -0.4239, 0.0000],
        [1.9857, -0.1365, 0.0000],
    ]
)


def ethanol_batch(n_struct: int, seed: int = 0, noise: float = 0.05):
    """Gas-phase ethanol-shaped molecules (BASELINE config 2): [B,9,3] positions."""
    rng = np.random.default_rng(seed)
    R = ETHANOL_R0[None] + rng.normal(scale=noise, size=(n_struct, 9, 3))
    return R, np.tile(ETHANOL_Z, (n_struct, 1))


def gmnn_params(seed: int = 1234, n_species: int = 119, n_radial: int = 5, n_basis: int = 7,
                units=(256, 256), n_out: int = 1, n_features: int = 360,
                scale: float | np.ndarray = 1.0, shift: float | np.ndarray = 0.0,
                random_scale_shift: bool = False, random_bias: bool = False):
    """Explicit random-init weight arrays in apax's parameter-tree layout.

    embedding ~ U(-1,1) (apax/layers/descriptor/basis_functions.py:101-120), dense ~
    truncated normal with std sqrt(1/fan_in) ("lecun"), biases 0, scale 1, shift 0.
    JAX's PRNG stream cannot be reproduced, so parity is always on identical explicit
    arrays.
    """
    rng = np.random.default_rng(seed)
    emb = rng.uniform(-1.0, 1.0, size=(n_species, n_species, n_radial, n_basis)).astype(np.float32)
    dims = [n_features] + list(units) + [n_out]
    readout = {}
    for li in range(len(dims) - 1):
        fan_in, fan_out = dims[li], dims[li + 1]
        std = np.sqrt(1.0 / fan_in) / 0.87962566103423978  # lecun_normal truncated-normal correction
        w = np.clip(rng.normal(size=(fan_in, fan_out)), -2.0, 2.0) * std
        b = rng.normal(size=(fan_out,)) * 0.1 if random_bias else np.zeros((fan_out,))
        readout[f"dense_{li}"] = {"w": w.astype(np.float32), "b": b.astype(np.float32)}
    if random_scale_shift:
        sc = rng.uniform(0.5, 1.5, size=(n_species, 1))
        sh = rng.normal(size=(n_species, 1))
    else:
        sc = np.broadcast_to(np.asarray(scale, dtype=np.float64).reshape(-1, 1), (n_species, 1)).copy()
        sh = np.broadcast_to(np.asarray(shift, dtype=np.float64).reshape(-1, 1), (n_species, 1)).copy()
    return {
        "params": {
            "energy_model": {
                "representation": {"radial_fn": {"atomic_type_embedding": emb}},
                "readout": readout,
                "scale_shift": {
                    "scale_per_element": sc.astype(np.float64),
                    "shift_per_element": sh.astype(np.float64),
                },
            }
        }
    }
