"""Host mirror of apax/utils/jax_md_reduced/partition.py's neighbour-list provider:
neighbor_list(...) -> NeighborListFns(allocate, update); NeighborList carries idx [2, capacity] int32
(device), reference_position, an error code with the same bits, and `.update(position, box=...)`.
Lists are built by the CUDA cell-list kernels (apax_nl_build_minimage) -- the reference enumerates all
N^2 candidates (partition.py:580-584, 693-699)."""
from __future__ import annotations

from dataclasses import dataclass, replace
from enum import Enum, IntEnum
from typing import Any, Callable, Optional

import torch

from ... import runtime as rt


class PartitionErrorCode(IntEnum):
    """partition.py:168-191."""
    NONE = 0
    NEIGHBOR_LIST_OVERFLOW = 1 << 0
    CELL_LIST_OVERFLOW = 1 << 1
    CELL_SIZE_TOO_SMALL = 1 << 2
    MALFORMED_BOX = 1 << 3


PEC = PartitionErrorCode


class NeighborListFormat(Enum):
    Dense = 0
    Sparse = 1
    OrderedSparse = 2


Sparse = NeighborListFormat.Sparse


@dataclass
class PartitionError:
    code: torch.Tensor  # u8 [1] on the device


@dataclass
class NeighborList:
    """partition.py:234-306."""
    idx: torch.Tensor
    reference_position: torch.Tensor
    error: PartitionError
    cell_list_capacity: Optional[int]
    max_occupancy: int
    format: NeighborListFormat
    cell_size: Optional[float]
    cell_list_fn: Any
    update_fn: Callable

    def update(self, position, **kwargs) -> "NeighborList":
        return self.update_fn(position, self, **kwargs)

    @property
    def did_buffer_overflow(self) -> bool:
        return bool(int(self.error.code.item()) & (PEC.NEIGHBOR_LIST_OVERFLOW | PEC.CELL_LIST_OVERFLOW))


@dataclass
class NeighborListFns:
    allocate: Callable
    update: Callable

    def __iter__(self):
        return iter((self.allocate, self.update))


def neighbor_list(displacement_or_metric, box, r_cutoff: float, dr_threshold: float = 0.0,
                  capacity_multiplier: float = 1.25, disable_cell_list: bool = False, mask_self: bool = True,
                  custom_mask_function=None, fractional_coordinates: bool = False,
                  format: NeighborListFormat = NeighborListFormat.Sparse, **static_kwargs) -> NeighborListFns:
    """Same signature as partition.neighbor_list (partition.py:478-490).  Supported: the configuration
    apax uses (ase_calc.py:54-62, simulate.py:506-514): Sparse, fractional coordinates, mask_self."""
    if format is not NeighborListFormat.Sparse or not fractional_coordinates or not mask_self or custom_mask_function:
        raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "neighbor_list", "only Sparse + fractional + mask_self")
    cutoff = float(r_cutoff) + float(dr_threshold)
    box0 = rt.to_device(box, torch.float64)

    def build(position, capacity, err_code, box_d):
        idx, n_found, overflow = rt.nl_build_minimage(position, box_d, cutoff, capacity)
        if err_code is not None:
            overflow |= err_code  # errors are sticky across updates (partition.py:734)
        return idx, n_found, overflow

    def allocate_fn(position, extra_capacity: int = 0, **kwargs):
        position = rt.to_device(position, torch.float64)
        box_d = rt.to_device(kwargs["box"], torch.float64) if "box" in kwargs else box0
        n = int(position.shape[0])
        # not jit-able in the reference either (partition.py:491-496): occupancy is read back to size the buffer
        probe = max(1024, n * 64)
        while True:
            idx, n_found, overflow = build(position, probe, None, box_d)
            occupancy = int(n_found.item())
            if occupancy <= probe:
                break
            probe = occupancy
        max_occ = int(occupancy * capacity_multiplier + n * extra_capacity)   # partition.py:714-718
        max_occ = min(max_occ, n * (n - 1))
        if max_occ != probe:
            idx, n_found, overflow = build(position, max_occ, None, box_d)
        return NeighborList(idx, position.clone(), PartitionError(overflow), None, max_occ, format, 0, None, update_fn)

    def update_fn(position, neighbors: NeighborList, **kwargs):
        position = rt.to_device(position, torch.float64)
        box_d = rt.to_device(kwargs["box"], torch.float64) if "box" in kwargs else box0
        # lax.cond(any(d > (dr_threshold/2)^2), rebuild, keep) (partition.py:771-779); the 4-byte flag is read back
        flag = rt.nl_needs_rebuild(position, neighbors.reference_position, box_d, dr_threshold)
        if int(flag.item()) == 0:
            return neighbors
        idx, n_found, overflow = build(position, neighbors.max_occupancy, neighbors.error.code, box_d)
        return replace(neighbors, idx=idx, reference_position=position.clone(), error=PartitionError(overflow))

    return NeighborListFns(allocate_fn, update_fn)
