"""Host mirror of apax/nn/builder.py: GMNNBuilder turns the `model` section of an apax config
(apax/config/model_config.py) into the CUDA-backed model objects.  This is the plug-in seam a
maintainer would switch: `Builder = GMNNBuilder` (apax/config/model_config.py get_builder)."""
from __future__ import annotations

import numpy as np

from ..layers.descriptor.gaussian_moment_descriptor import BesselBasis, GaussianBasis, GaussianMomentDescriptor, RadialFunction
from .models import (AtomisticReadout, EnergyDerivativeModel, EnergyModel, ExponentialRepulsion, PerElementScaleShift,
                     ShallowEnsembleModel, ZBLRepulsion)

DEFAULT_MODEL_CONFIG = {
    "name": "gmnn", "basis": {"name": "gaussian", "n_basis": 7, "r_min": 0.5, "r_max": 6.0, "spacing": "linear"},
    "n_radial": 5, "n_contr": 8, "emb_init": "uniform", "nn": [256, 256], "w_init": "lecun", "b_init": "zeros",
    "use_ntk": False, "activation_fn": "variance_preserving_swish", "ensemble": None, "calc_stress": False,
    "property_heads": [], "empirical_corrections": [], "descriptor_dtype": "fp32", "readout_dtype": "fp32",
    "scale_shift_dtype": "fp64",
}


class ModelBuilder:
    def __init__(self, model_config: dict | None = None, n_species: int = 119):
        cfg = dict(DEFAULT_MODEL_CONFIG)
        cfg.update(model_config or {})
        self.config = cfg
        self.n_species = n_species

    def build_basis_function(self):
        b = self.config["basis"]
        if b["name"] == "bessel":                            # apax/nn/builder.py:36-52
            return BesselBasis(n_basis=b["n_basis"], r_max=b["r_max"], dtype=self.config["descriptor_dtype"])
        return GaussianBasis(n_basis=b["n_basis"], r_min=b["r_min"], r_max=b["r_max"], dtype=self.config["descriptor_dtype"],
                             spacing=b.get("spacing", "linear"))

    def build_radial_function(self):
        return RadialFunction(n_radial=self.config["n_radial"], basis_fn=self.build_basis_function(), n_species=self.n_species,
                              emb_init=self.config["emb_init"], use_embed_norm=True, one_sided_dist=False,
                              dtype=self.config["descriptor_dtype"])

    def build_descriptor(self, apply_mask):
        raise NotImplementedError("use a subclass to facilitate this")

    def build_readout(self, head_config, is_feature_fn=False, only_use_n_layers=None):
        ens = head_config.get("ensemble")
        n_shallow = ens["n_members"] if ens and ens.get("kind") == "shallow" else 0
        return AtomisticReadout(units=list(head_config["nn"]), b_init=head_config["b_init"], w_init=head_config["w_init"],
                                use_ntk=head_config["use_ntk"], n_shallow_ensemble=n_shallow,
                                dtype=head_config.get("readout_dtype", "fp32"), activation_fn=self.config["activation_fn"])

    def build_corrections(self, apply_mask: bool = True):
        """apax/nn/builder.py:146-158: [{"name": "zbl" | "exponential", "r_max": ...}, ...]."""
        out = []
        for c in self.config["empirical_corrections"]:
            c = dict(c)
            name = c.pop("name")
            if name == "zbl":
                out.append(ZBLRepulsion(apply_mask=apply_mask, **c))
            elif name == "exponential":
                out.append(ExponentialRepulsion(apply_mask=apply_mask, **c))
            else:
                raise NotImplementedError(f"empirical correction '{name}' is not on the CUDA path")
        return out

    def build_scale_shift(self, scale, shift):
        return PerElementScaleShift(n_species=self.n_species, scale=scale, shift=shift, dtype=self.config["scale_shift_dtype"])

    def build_energy_model(self, scale=1.0, shift=0.0, apply_mask=True, init_box=np.array([0.0, 0.0, 0.0]),
                           inference_disp_fn=None):
        return EnergyModel(representation=self.build_descriptor(apply_mask), readout=self.build_readout(self.config),
                           scale_shift=self.build_scale_shift(scale, shift), property_heads=list(self.config["property_heads"]),
                           corrections=self.build_corrections(apply_mask=apply_mask), init_box=init_box,
                           inference_disp_fn=inference_disp_fn)

    def build_energy_derivative_model(self, scale=1.0, shift=0.0, apply_mask=True, init_box=np.array([0.0, 0.0, 0.0]),
                                      inference_disp_fn=None):
        em = self.build_energy_model(scale, shift, apply_mask, init_box=init_box, inference_disp_fn=inference_disp_fn)
        ens = self.config["ensemble"]
        if ens and ens.get("kind") == "shallow" and ens.get("n_members", 0) > 1:
            return ShallowEnsembleModel(em, calc_stress=self.config["calc_stress"], force_variance=ens.get("force_variance", True),
                                        chunk_size=ens.get("chunk_size"))
        return EnergyDerivativeModel(em, calc_stress=self.config["calc_stress"])


class GMNNBuilder(ModelBuilder):
    def build_descriptor(self, apply_mask):
        return GaussianMomentDescriptor(radial_fn=self.build_radial_function(), n_contr=self.config["n_contr"],
                                        dtype=self.config["descriptor_dtype"], apply_mask=apply_mask)
