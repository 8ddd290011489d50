"""JAX side of the XLA-FFI layer: `jax.ffi.ffi_call` wrappers over the handlers of xla_ffi_shim.cc and the
`jax.custom_vjp` rules that make them differentiable where apax differentiates them.

IMPORT-GATED: jax / jaxlib are not in this repository's build image nor on its GPU boxes, so nothing here
has been executed; the module imports without jax and every entry point raises `FfiUnavailable` until jax is importable.
It is the code a maintainer drops next to apax to switch the hot path, kept mechanical on purpose: each function only
shapes buffers and forwards them; all arithmetic is in libapax_b200.so.  The tested front-end of the same C ABI in this
repository is the ctypes runtime (apax_b200/runtime.py); the handler bodies are kept under the compiler's eye by
tests/test_host_logic.py (syntax check against a stub of the FFI types).

What replaces what (reference path:line):
  energy_fn / energy_and_forces  EnergyModel.apply + jax.grad / EnergyDerivativeModel.apply   apax/nn/models.py:87-171,
                                 differentiated at apax/utils/jax_md_reduced/quantity.py:52-54 (canonicalize_force)
  energy_forces_stress           EnergyDerivativeModel(calc_stress=True)                      apax/nn/models.py:160-169
  descriptor                     GaussianMomentDescriptor.__call__ behind build_descriptor    apax/nn/builder.py:79-83, 248-260
  neighbor_list_minimage,
  needs_rebuild                  partition.neighbor_list allocate / update                    apax/utils/jax_md_reduced/partition.py:478-787
  train_loss_and_grad            calc_loss + jax.value_and_grad w.r.t. params                 apax/train/trainer.py:253-263, 332-336
"""
from __future__ import annotations

import ctypes as C
from .build_ffi import FFI_TARGETS, FfiUnavailable, register as _register_targets

DISP = {"gas": 0, "offsets": 1, "minimage": 2, "drvec": 3}
N_FEATURES = 360
_registered = False


def _jax():
    try:
        import jax
        import jax.numpy as jnp
    except Exception as exc:  # pragma: no cover - depends on the image
        raise FfiUnavailable(f"jax is not importable here ({type(exc).__name__}: {exc})") from exc
    return jax, jnp


def register(lib_path: str | None = None) -> None:
    """Build (if needed) and register every handler of the shim: FFI_TARGETS, platform "CUDA".  Idempotent."""
    global _registered
    _jax()
    if not _registered:
        _register_targets(lib_path)
        _registered = True


def _lib():
    from .. import _lib as L

    return L.lib()


def _handle(device_params) -> int:
    """The opaque apax_gm_params* of a runtime.DeviceParams (or a raw integer handle) as the int64 attribute the handlers take."""
    h = getattr(device_params, "handle", device_params)
    return int(getattr(h, "value", h))


def gm_workspace_bytes(n_atoms: int, n_pairs: int, n_out: int = 1) -> int:
    return int(_lib().apax_gm_workspace_bytes(C.c_int64(n_atoms), C.c_int64(n_pairs), C.c_int32(n_out)))


def nl_workspace_bytes(n_atoms: int, capacity: int) -> int:
    return int(_lib().apax_nl_workspace_bytes(C.c_int64(n_atoms), C.c_int64(capacity)))


def _empty_f64(jnp):
    return jnp.zeros((0,), dtype=jnp.float64)


# ------------------------------------------------------------------------------------------- energy + forces
def energy_and_forces(device_params, R, Z, idx, box, offsets=None, mode: str = "minimage", n_out: int = 1):
    """One fused evaluation: (energy f64 [n_out], forces f64 [n_out, N, 3]).  Shapes are static under jit (N, P from the
    arguments), the scratch is a third result sized on the host."""
    jax, jnp = _jax()
    register()
    n, p = R.shape[0], idx.shape[1]
    out = (jax.ShapeDtypeStruct((n_out,), jnp.float64), jax.ShapeDtypeStruct((n_out, n, 3), jnp.float64),
           jax.ShapeDtypeStruct((gm_workspace_bytes(n, p, n_out),), jnp.uint8))
    call = jax.ffi.ffi_call("apax_energy_forces", out, vmap_method="sequential")
    e, f, _ = call(R.astype(jnp.float64), Z.astype(jnp.int32), idx.astype(jnp.int32), jnp.asarray(box, jnp.float64).reshape(3, 3),
                   _empty_f64(jnp) if offsets is None else offsets.astype(jnp.float64),
                   params=_handle(device_params), disp_mode=DISP[mode])
    return e, f


def make_energy_fn(device_params, Z, mode: str = "minimage"):
    """The callable JaxMD integrators differentiate (apax/md/simulate.py:41-74, create_energy_fn): scalar energy of
    `R` with a custom VJP whose backward rule is the force array the SAME fused call already produced --
    jax.grad(energy_fn) (quantity.py:52-54) therefore costs no second evaluation.

        energy_fn(R, neighbor, box, offsets=None) -> f64 scalar        (mean over heads for a shallow ensemble)
    """
    jax, jnp = _jax()

    @jax.custom_vjp
    def _energy(R, idx, box, offsets):
        e, _ = energy_and_forces(device_params, R, Z, idx, box, offsets, mode, _n_out(device_params))
        return jnp.mean(e)

    def _fwd(R, idx, box, offsets):
        e, f = energy_and_forces(device_params, R, Z, idx, box, offsets, mode, _n_out(device_params))
        return jnp.mean(e), (jnp.mean(f, axis=0), idx, box, offsets)

    def _bwd(res, ct):
        f_mean, idx, box, offsets = res
        # dE/dR = -F (Cartesian also for fractional R: identity Jacobian of space.transform, space.py:93-97);
        # the list is integer data, box / offsets carry no gradient on the E+F path
        zero = lambda x: None if x is None else jnp.zeros_like(x)  # noqa: E731
        return (-ct * f_mean, None, zero(box), zero(offsets))

    _energy.defvjp(_fwd, _bwd)

    def energy_fn(R, neighbor, box, offsets=None, **unused):
        idx = getattr(neighbor, "idx", neighbor)
        return _energy(R, idx, box, offsets)

    return energy_fn


def _n_out(device_params) -> int:
    return int(getattr(device_params, "n_out", 1))


def energy_forces_stress(device_params, R, Z, idx, box, offsets=None, mode: str = "minimage"):
    """EnergyDerivativeModel with calc_stress (apax/nn/models.py:160-169): (energy [1], forces [1,N,3], stress x volume [3,3])."""
    jax, jnp = _jax()
    register()
    n, p = R.shape[0], idx.shape[1]
    f64 = jnp.float64
    out = (jax.ShapeDtypeStruct((1,), f64), jax.ShapeDtypeStruct((1, n, 3), f64), jax.ShapeDtypeStruct((3, 3), f64),
           jax.ShapeDtypeStruct((1,), f64), jax.ShapeDtypeStruct((1, n, 3), f64),
           jax.ShapeDtypeStruct((gm_workspace_bytes(n, p, 1),), jnp.uint8))
    box = jnp.asarray(box, f64).reshape(3, 3)
    box2 = 2.0 * box if mode == "offsets" else _empty_f64(jnp)       # the offsets branch differentiates at doubled displacements
    e, f, s, _, _, _ = jax.ffi.ffi_call("apax_energy_forces_stress", out)(
        R.astype(f64), Z.astype(jnp.int32), idx.astype(jnp.int32), box, _empty_f64(jnp) if offsets is None else offsets.astype(f64),
        box2, params=_handle(device_params), disp_mode=DISP[mode])
    return e, f, s


# ------------------------------------------------------------------------------------------- descriptor seam
def make_descriptor(device_params, n_features: int = N_FEATURES):
    """`representation(dr_vec [P,3] f64, Z [N], idx [2,P]) -> g [N, n_features] f32` (apax/nn/models.py:107) as a
    jax.custom_vjp function: forward = apax_gm_descriptor, backward = apax_gm_descriptor_vjp (both APAX_DISP_DRVEC).
    A Flax module that calls this in __call__ is a drop-in for GaussianMomentDescriptor behind
    ModelBuilder.build_descriptor; the readout, scale/shift and jax.grad of the host stay untouched."""
    jax, jnp = _jax()
    handle = _handle(device_params)

    def _fwd_call(dr_vec, Z, idx):
        register()
        n, p = Z.shape[0], idx.shape[1]
        out = (jax.ShapeDtypeStruct((n, N_FEATURES), jnp.float32), jax.ShapeDtypeStruct((gm_workspace_bytes(n, p, 1),), jnp.uint8))
        g, _ = jax.ffi.ffi_call("apax_descriptor", out)(dr_vec.astype(jnp.float64), Z.astype(jnp.int32), idx.astype(jnp.int32),
                                                        _empty_f64(jnp), _empty_f64(jnp), params=handle, disp_mode=DISP["drvec"])
        return g[:, :n_features]                                          # n_contr truncation (gaussian_moment_descriptor.py:101)

    @jax.custom_vjp
    def descriptor(dr_vec, Z, idx):
        return _fwd_call(dr_vec, Z, idx)

    def fwd(dr_vec, Z, idx):
        return _fwd_call(dr_vec, Z, idx), (dr_vec, Z, idx)

    def bwd(res, g_bar):
        dr_vec, Z, idx = res
        n, p = Z.shape[0], idx.shape[1]
        gb = jnp.zeros((n, N_FEATURES), jnp.float32).at[:, :n_features].set(g_bar.astype(jnp.float32))
        out = (jax.ShapeDtypeStruct((p, 3), jnp.float64), jax.ShapeDtypeStruct((0,), jnp.float64),
               jax.ShapeDtypeStruct((gm_workspace_bytes(n, p, 1),), jnp.uint8))
        d_dr, _, _ = jax.ffi.ffi_call("apax_descriptor_vjp", out)(dr_vec.astype(jnp.float64), Z.astype(jnp.int32), idx.astype(jnp.int32),
                                                                  _empty_f64(jnp), _empty_f64(jnp), gb, params=handle,
                                                                  disp_mode=DISP["drvec"])
        return (d_dr.astype(dr_vec.dtype), None, None)

    descriptor.defvjp(fwd, bwd)
    return descriptor


# ------------------------------------------------------------------------------------------- neighbour list
def neighbor_list_minimage(frac, box, cutoff: float, capacity: int):
    """partition.neighbor_list(...).allocate / update body (partition.py:640-664): idx i32 [2, capacity] in the reference's
    order, padded with N; n_found i32 [1]; overflow u8 [1] (PEC.NEIGHBOR_LIST_OVERFLOW)."""
    jax, jnp = _jax()
    register()
    n = frac.shape[0]
    out = (jax.ShapeDtypeStruct((2, capacity), jnp.int32), jax.ShapeDtypeStruct((1,), jnp.int32), jax.ShapeDtypeStruct((1,), jnp.uint8),
           jax.ShapeDtypeStruct((nl_workspace_bytes(n, capacity),), jnp.uint8))
    idx, n_found, overflow, _ = jax.ffi.ffi_call("apax_nl_minimage", out)(frac.astype(jnp.float64),
                                                                        jnp.asarray(box, jnp.float64).reshape(3, 3), cutoff=float(cutoff))
    return idx, n_found, overflow


def needs_rebuild(frac, reference_frac, box, dr_threshold: float):
    """The skin test (partition.py:771-779): i32 [1], non-zero when any atom moved further than dr_threshold / 2."""
    jax, jnp = _jax()
    register()
    out = jax.ShapeDtypeStruct((1,), jnp.int32)
    return jax.ffi.ffi_call("apax_nl_needs_rebuild", out)(frac.astype(jnp.float64), reference_frac.astype(jnp.float64),
                                                           jnp.asarray(box, jnp.float64).reshape(3, 3), dr_threshold=float(dr_threshold))


# ------------------------------------------------------------------------------------------- training step
def train_loss_and_grad(plan_handle: int, theta32, theta64, batch: dict, loss_cfg: dict, batch_divisor: float, workspace_bytes: int):
    """calc_loss + jax.value_and_grad w.r.t. the parameters (apax/train/trainer.py:253-263, 332-336) for the flat parameter
    blocks of apax_train_param_layout: returns (loss [1], energy [B], forces [N,3], grad32, grad64).  `batch` holds R, Z, idx,
    offsets (or None), struct_ptr, pair_ptr, n_real, E_label, F_label in the layout of apax_gm_energy_forces_batched."""
    jax, jnp = _jax()
    register()
    f64, i32 = jnp.float64, jnp.int32
    n, b = batch["R"].shape[0], batch["E_label"].shape[0]
    out = (jax.ShapeDtypeStruct((1,), f64), jax.ShapeDtypeStruct((b,), f64), jax.ShapeDtypeStruct((n, 3), f64),
           jax.ShapeDtypeStruct(theta32.shape, jnp.float32), jax.ShapeDtypeStruct(theta64.shape, f64),
           jax.ShapeDtypeStruct((int(workspace_bytes),), jnp.uint8))
    off = batch.get("offsets")
    loss, energy, forces, g32, g64, _ = jax.ffi.ffi_call("apax_train_loss_and_grad", out)(
        theta32.astype(jnp.float32), theta64.astype(f64), batch["R"].astype(f64), batch["Z"].astype(i32), batch["idx"].astype(i32),
        _empty_f64(jnp) if off is None else off.astype(f64), batch["struct_ptr"].astype(i32), batch["pair_ptr"].astype(i32),
        batch["n_real"].astype(i32), batch["E_label"].astype(f64), batch["F_label"].astype(f64), plan=int(plan_handle),
        energy_weight=float(loss_cfg["energy_weight"]), energy_atoms_exponent=float(loss_cfg["energy_atoms_exponent"]),
        forces_weight=float(loss_cfg["forces_weight"]), forces_atoms_exponent=float(loss_cfg["forces_atoms_exponent"]),
        batch_divisor=float(batch_divisor))
    return loss, energy, forces, g32, g64


__all__ = ["FFI_TARGETS", "FfiUnavailable", "register", "energy_and_forces", "make_energy_fn", "energy_forces_stress",
           "make_descriptor", "neighbor_list_minimage", "needs_rebuild", "train_loss_and_grad", "gm_workspace_bytes",
           "nl_workspace_bytes"]
