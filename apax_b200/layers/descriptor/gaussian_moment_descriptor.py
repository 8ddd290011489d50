"""Host mirror of apax/layers/descriptor/gaussian_moment_descriptor.py behind the builder seam
(apax/nn/builder.py:79-83, 248-260): same constructor fields, same call contract
`representation(dr_vec [P,3], Z [N], neighbor_idxs [2,P]) -> g [N, 360]`, computed by the fused CUDA
pair + contraction kernels (apax_gm_descriptor with APAX_DISP_DRVEC)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

from ... import runtime as rt


@dataclass
class GaussianBasis:
    """apax/layers/descriptor/basis_functions.py:12-17 (statics only; evaluated inside the pair kernels)."""
    n_basis: int = 7
    r_min: float = 0.5
    r_max: float = 6.0
    dtype: Any = "fp32"
    spacing: str = "linear"


@dataclass
class BesselBasis:
    """apax/layers/descriptor/basis_functions.py:59-79 (statics only)."""
    n_basis: int = 7
    r_max: float = 6.0
    dtype: Any = "fp32"


def basis_kind(basis_fn) -> int:
    """apax_gm_config.basis_kind: 0 Gaussian linear, 1 Gaussian exponential, 2 Bessel."""
    if isinstance(basis_fn, BesselBasis):
        return 2
    if basis_fn.spacing == "linear":
        return 0
    if basis_fn.spacing == "exponential":
        return 1
    raise NotImplementedError(f"spacing {basis_fn.spacing} has not been implemented. Available options are: ['linear', 'exponential']")


@dataclass
class RadialFunction:
    """apax/layers/descriptor/basis_functions.py:88-95 (statics only)."""
    n_radial: int = 5
    basis_fn: GaussianBasis = field(default_factory=GaussianBasis)
    n_species: int = 119
    emb_init: str = "uniform"
    use_embed_norm: bool = True
    one_sided_dist: bool = False
    dtype: Any = "fp32"


@dataclass
class GaussianMomentDescriptor:
    radial_fn: RadialFunction = field(default_factory=RadialFunction)
    n_contr: int = 8
    dtype: Any = "fp32"
    apply_mask: bool = True

    def statics(self) -> dict:
        b = self.radial_fn.basis_fn
        if self.radial_fn.emb_init != "uniform" or self.dtype not in ("fp32", None) or not self.apply_mask:
            raise rt._lib.ApaxB200Error(rt._lib.ERR_UNSUPPORTED, "GaussianMomentDescriptor",
                                        "kernels cover the embedded radial function, fp32, apply_mask=True")
        return dict(n_basis=b.n_basis, r_min=getattr(b, "r_min", 0.0), r_max=b.r_max, n_radial=self.radial_fn.n_radial,
                    n_species=self.radial_fn.n_species, use_embed_norm=self.radial_fn.use_embed_norm, n_contr=self.n_contr,
                    basis_kind=basis_kind(b))

    def apply(self, variables: dict, dr_vec, Z, neighbor_idxs):
        """variables = {"params": {"radial_fn": {"atomic_type_embedding": [S,S,5,7]}}} as flax would pass it."""
        emb = variables["params"]["radial_fn"]["atomic_type_embedding"]
        st = self.statics()
        dp = _descriptor_params(emb, st)
        return rt.descriptor(dp, dr_vec, Z, neighbor_idxs, mode="drvec")

    def vjp(self, variables: dict, dr_vec, Z, neighbor_idxs, g_bar):
        """Backward rule of `apply` with respect to dr_vec (apax_gm_descriptor_vjp): g_bar [N, n_feat] -> dr_vec_bar [P,3] f64.
        This is the pair (fwd, bwd) a JAX host registers as jax.custom_vjp around the descriptor seam
        (apax_b200/ffi/jax_binding.py::descriptor)."""
        emb = variables["params"]["radial_fn"]["atomic_type_embedding"]
        st = self.statics()
        dp = _descriptor_params(emb, st)
        return rt.descriptor_vjp(dp, dr_vec, Z, neighbor_idxs, g_bar, mode="drvec")[0]


_cache: dict = {}


def _descriptor_params(emb, statics):
    import numpy as np

    key = (id(emb), tuple(sorted(statics.items())))
    if key not in _cache:
        zeros = lambda *s: np.zeros(s, dtype=np.float32)  # noqa: E731
        tree = {"representation": {"radial_fn": {"atomic_type_embedding": emb}},
                "readout": {"dense_0": {"w": zeros(rt.FEATURE_COUNTS[statics["n_contr"] - 1], 256), "b": zeros(256)}, "dense_1": {"w": zeros(256, 256), "b": zeros(256)},
                            "dense_2": {"w": zeros(256, 1), "b": zeros(1)}},
                "scale_shift": {"scale_per_element": np.ones((statics["n_species"], 1)),
                                "shift_per_element": np.zeros((statics["n_species"], 1))}}
        cfg = {k: statics[k] for k in ("n_basis", "r_min", "r_max", "n_radial", "use_embed_norm", "n_contr", "basis_kind")}
        _cache[key] = rt.DeviceParams(tree, cfg)
    return _cache[key]
