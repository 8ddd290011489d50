"""Compile the XLA-FFI shim (xla_ffi_shim.cc) against jaxlib's headers -- possible only where JAX is installed, which
is neither this repository's build image nor its GPU boxes.  `build_ffi()` returns the path of
libapax_b200_ffi.so, or raises FfiUnavailable with the reason; nothing else in the package depends on it."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class FfiUnavailable(RuntimeError):
    pass


def build_ffi() -> str:
    try:
        import jax  # noqa: F401
        from jax import ffi as jffi
    except Exception as exc:  # pragma: no cover - depends on the image
        raise FfiUnavailable(f"jax is not importable here ({type(exc).__name__}: {exc}); the shim needs xla/ffi/api/ffi.h") from exc
    inc = jffi.include_dir()
    cxx = shutil.which("g++") or shutil.which("c++")
    if not cxx:
        raise FfiUnavailable("no C++ compiler")
    out = os.path.join(ROOT, "libapax_b200_ffi.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{inc}", f"-I{cuda_inc}", f"-I{os.path.join(os.path.dirname(ROOT), 'include')}",
           os.path.join(HERE, "xla_ffi_shim.cc"), f"-L{ROOT}", "-l:libapax_b200.so", f"-Wl,-rpath,{ROOT}", "-o", out]
    subprocess.check_call(cmd)
    return out


# XLA-FFI target name -> handler symbol of xla_ffi_shim.cc: every handler the shim defines
FFI_TARGETS = (
    ("apax_energy_forces", "ApaxEnergyForces"),
    ("apax_energy_forces_stress", "ApaxEnergyForcesStress"),
    ("apax_descriptor", "ApaxDescriptor"),
    ("apax_descriptor_vjp", "ApaxDescriptorVjp"),
    ("apax_nl_minimage", "ApaxNeighborListMinImage"),
    ("apax_nl_needs_rebuild", "ApaxNeedsRebuild"),
    ("apax_train_loss_and_grad", "ApaxTrainLossAndGrad"),
)


def register(lib_path: str | None = None) -> None:
    """jax.ffi.register_ffi_target for every handler of the shim (platform "CUDA")."""
    import ctypes

    import jax

    lib = ctypes.CDLL(lib_path or build_ffi())
    for target, symbol in FFI_TARGETS:
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
