"""Host mirrors of apax/nn/models.py: EnergyModel (:70-134), EnergyDerivativeModel (:137-171) and
ShallowEnsembleModel (:199-275) with the same `.apply(params, R, Z, neighbor, box, offsets)` call
shape.  The whole of EnergyModel.__call__ and its reverse mode is one C-ABI call
(apax_gm_energy_forces); nothing is computed in Python."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np
import torch

from .. import runtime as rt
from ..layers.descriptor.gaussian_moment_descriptor import GaussianMomentDescriptor


@dataclass
class AtomisticReadout:
    """apax/layers/readout.py:12-20 (statics only)."""
    units: list = field(default_factory=lambda: [256, 256])
    activation_fn: Any = "variance_preserving_swish"
    w_init: str = "lecun"
    b_init: str = "zeros"
    use_ntk: bool = False
    n_shallow_ensemble: int = 0
    dtype: Any = "fp32"


@dataclass
class ZBLRepulsion:
    """apax/layers/empirical.py:25-27 (statics only; parameters live in the tree under corrections_<i>)."""
    r_max: float = 2.0
    apply_mask: bool = True
    dtype: Any = "fp32"


@dataclass
class ExponentialRepulsion:
    """apax/layers/empirical.py:89-91 (statics only)."""
    r_max: float = 2.0
    apply_mask: bool = True
    dtype: Any = "fp32"


@dataclass
class PerElementScaleShift:
    """apax/layers/scaling.py:10-14 (statics only; values live in the parameter tree)."""
    n_species: int = 119
    scale: Any = 1.0
    shift: Any = 0.0
    dtype: Any = "fp64"


def canonicalize_neighbors(neighbor):
    """apax/layers/distances.py:8-9."""
    return neighbor.idx if hasattr(neighbor, "idx") else neighbor


def displacement_mode(init_box, inference_disp_fn) -> str:
    """The three branches of make_distance_fn (apax/layers/distances.py:32-46)."""
    if np.all(np.asarray(init_box) < 1e-6):
        return "gas"
    return "offsets" if inference_disp_fn is None else "minimage"


_param_cache: dict = {}


def device_params(params: dict, config: dict) -> rt.DeviceParams:
    """Upload a parameter tree once per (tree object, statics)."""
    key = (id(params), tuple(sorted((k, str(v)) for k, v in config.items())))
    hit = _param_cache.get(key)
    if hit is None or hit[0] is not params:
        _param_cache[key] = (params, rt.DeviceParams(params, config))
    return _param_cache[key][1]


@dataclass
class EnergyModel:
    representation: GaussianMomentDescriptor = field(default_factory=GaussianMomentDescriptor)
    readout: AtomisticReadout = field(default_factory=AtomisticReadout)
    scale_shift: PerElementScaleShift = field(default_factory=PerElementScaleShift)
    property_heads: list = field(default_factory=list)
    corrections: list = field(default_factory=list)
    init_box: Any = field(default_factory=lambda: np.array([0.0, 0.0, 0.0]))
    mask_atoms: bool = True
    inference_disp_fn: Any = None

    def statics(self) -> dict:
        if self.property_heads:
            raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "EnergyModel", "property heads")
        corr = {}
        for c in self.corrections:
            if isinstance(c, ZBLRepulsion):
                corr["zbl_r_max"] = float(c.r_max)
            elif isinstance(c, ExponentialRepulsion):
                corr["exp_r_max"] = float(c.r_max)
            else:
                raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "EnergyModel", f"empirical correction {type(c).__name__}")
        if list(self.readout.units) != [256, 256] or self.readout.activation_fn != "variance_preserving_swish":
            raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "EnergyModel", "readout must be [256,256] + vp-swish")
        cfg = self.representation.statics()
        cfg.update(nn=list(self.readout.units), use_ntk=self.readout.use_ntk, mask_atoms=self.mask_atoms)
        cfg.update(corr)
        return cfg

    def _run(self, params, R, Z, neighbor, box, offsets, forces: bool):
        mode = displacement_mode(self.init_box, self.inference_disp_fn)
        st = self.statics()
        # every radial basis (7 Gaussians, Bessel, exponential spacing, n_basis <= 16) runs on the fused pair kernels and the
        # integer tensor-core readout; the table-driven kernels behind apax_train_* serve training only
        dp = device_params(params, st)
        return rt.energy_forces(dp, R, Z, canonicalize_neighbors(neighbor), box, offsets, mode, forces=forces)

    def apply(self, params, R, Z, neighbor, box, offsets, perturbation=None):
        """-> (energy f64 scalar tensor, or [n_ens] for a shallow ensemble; properties {}) (models.py:87-134)."""
        if perturbation is not None:
            raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "EnergyModel", "strain perturbation (stress) is a next row")
        e, _ = self._run(params, R, Z, neighbor, box, offsets, forces=False)
        return (e[0] if e.numel() == 1 else e), {}


@dataclass
class EnergyDerivativeModel:
    energy_model: EnergyModel = field(default_factory=EnergyModel)
    calc_stress: bool = False

    def apply(self, params, R, Z, neighbor, box, offsets):
        """-> {"energy": f64 scalar, "forces": [N,3] f64} (models.py:146-171)."""
        if self.calc_stress:
            # stress_times_vol (apax/layers/properties.py:11-42).  Min-image path: strain applied as (I + eps) . dr
            # (space.py:277-278); offsets path: as the reference's disp_fn does it, dR + (I + eps) . dR (+ offsets), i.e. at
            # doubled displacements (apax/layers/distances.py:12-22) -- reproduced as written.
            em = self.energy_model
            mode = displacement_mode(em.init_box, em.inference_disp_fn)
            st = em.statics()
            if mode not in ("minimage", "offsets"):
                raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "EnergyDerivativeModel", "calc_stress needs a periodic displacement")
            e, f, s = rt.energy_forces(device_params(params, st), R, Z, canonicalize_neighbors(neighbor), box, offsets, mode,
                                       stress=True)
            return {"energy": e[0], "forces": f[0], "stress": s}
        e, f = self.energy_model._run(params, R, Z, neighbor, box, offsets, forces=True)
        return {"energy": e[0], "forces": f[0]}


@dataclass
class ShallowEnsembleModel:
    energy_model: EnergyModel = field(default_factory=EnergyModel)
    calc_stress: bool = False
    force_variance: bool = True
    chunk_size: Optional[int] = None  # accepted for parity of signature; all heads are evaluated in one call

    def apply(self, params, R, Z, neighbor, box, offsets):
        """models.py:209-275: energy statistics over heads, per-head forces and their statistics."""
        if not self.force_variance:
            # forces_mean = -jax.grad(mean_energy_fn) (models.py:265-267): ONE backward pass with the mean head's output weights
            em = self.energy_model
            dp = device_params(params, {**em.statics(), "mean_forces": True})
            e_ens, f_mean = rt.energy_forces(dp, R, Z, canonicalize_neighbors(neighbor), box, offsets,
                                             displacement_mode(em.init_box, em.inference_disp_fn))
            n_ens = e_ens.shape[0]
            e_mean = e_ens.mean()
            return {"energy": e_mean, "energy_ensemble": e_ens,
                    "energy_uncertainty": torch.sqrt(torch.sum((e_ens - e_mean) ** 2) / (n_ens - 1)), "forces": f_mean[0]}
        e_ens, f_ens = self.energy_model._run(params, R, Z, neighbor, box, offsets, forces=True)
        n_ens = e_ens.shape[0]
        divisor = 1.0 / (n_ens - 1)
        e_mean = e_ens.mean()
        pred = {"energy": e_mean, "energy_ensemble": e_ens,
                "energy_uncertainty": torch.sqrt(divisor * torch.sum((e_ens - e_mean) ** 2))}
        f_mean = f_ens.mean(dim=0)
        pred["forces"] = f_mean
        if self.force_variance:
            pred["forces_uncertainty"] = torch.sqrt(divisor * torch.sum((f_ens - f_mean) ** 2, dim=0))
            pred["forces_ensemble"] = f_ens.permute(1, 2, 0)  # n_atoms x 3 x n_members (:262-263)
        return pred
