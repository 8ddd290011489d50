"""CPU-side checks: the generated contraction program, and that the C-ABI library loads and exports
every symbol include/apax_b200.h declares (no compute calls without a GPU)."""
import ctypes
import itertools
import os
import re

import numpy as np
import pytest
import torch

from apax_b200 import _lib
from apax_b200 import synthetic as syn
from apax_b200.csrc import gen_contract as gc
from oracle import gmnn_oracle as go
from oracle import neighbor_oracle as no

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pack(m0, m1, m2, m3):
    n = m0.shape[0]
    out = np.zeros((100, n))
    for r in range(5):
        out[r * 20] = m0[:, r]
        for i in range(3):
            out[r * 20 + 1 + i] = m1[:, r, i]
        for i, j in itertools.combinations_with_replacement(range(3), 2):
            out[r * 20 + 4 + gc.pack2(i, j)] = m2[:, r, i, j]
        for i, j, k in itertools.combinations_with_replacement(range(3), 3):
            out[r * 20 + 10 + gc.pack3(i, j, k)] = m3[:, r, i, j, k]
    return out


def test_generated_contraction_program_matches_oracle():
    blocks, nf = gc.build_program()
    assert nf == 360
    pos, Z, cell = syn.water_box(6, 3)
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
    aux = go.descriptor_only(syn.gmnn_params(), pos @ np.linalg.inv(cell), Z, idx, cell.T, off, "offsets", go.fp64_config())
    mp = _pack(*aux["moments"])
    g = gc.run_forward(blocks, mp)
    assert np.abs(g.T - aux["g"]).max() < 1e-12 * max(1.0, np.abs(aux["g"]).max())
    # reverse mode against a finite-difference of sum(gbar * g)
    rng = np.random.default_rng(0)
    gbar = rng.normal(size=g.shape)
    a = gc.run_backward(blocks, mp, gbar)
    for slot in (0, 3, 27, 55, 99):
        h = 1e-6
        mp_p, mp_m = mp.copy(), mp.copy()
        mp_p[slot] += h
        mp_m[slot] -= h
        fd = ((gc.run_forward(blocks, mp_p) - gc.run_forward(blocks, mp_m)) * gbar).sum(0) / (2 * h)
        assert np.abs(fd - a[slot]).max() < 1e-5 * max(1.0, np.abs(a[slot]).max())


def test_looped_reverse_program_equals_straight_line_program():
    """The CUDA backward kernel runs the looped form (generic block per contraction type + instance tables);
    expanded back to concrete blocks it must be the same program as the straight-line one."""
    blocks, nf = gc.build_program()
    types, off, nf2 = gc.build_looped()
    assert nf2 == nf == 360
    looped = gc.expand_looped(types, off)
    rng = np.random.default_rng(1)
    m = rng.normal(size=(100, 9))
    gbar = rng.normal(size=(nf, 9))
    assert np.abs(gc.run_forward(blocks, m) - gc.run_forward(looped, m)).max() < 1e-12
    a_ref, a_loop = gc.run_backward(blocks, m, gbar), gc.run_backward(looped, m, gbar)
    assert np.abs(a_ref - a_loop).max() < 1e-12 * np.abs(a_ref).max()
    # every feature is produced exactly once
    feats = sorted(f for ty in types for _, _, fs in ty["instances"] for f in fs if f >= 0)
    assert feats == list(range(5, 360))


def test_generated_header_is_current():
    path = os.path.join(ROOT, "apax_b200", "csrc", "contract_gen.cuh")
    tmp = path + ".check"
    gc.main(tmp)
    try:
        assert open(tmp).read() == open(path).read()
    finally:
        os.remove(tmp)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "apax_b200.h")).read()
    declared = set(re.findall(r"\b(apax_[a-z0-9_]+)\s*\(", header))
    declared -= {"apax_gm_params", "apax_ctx", "apax_md_nvt_config", "apax_gm_config"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libapax_b200.so not built in this checkout (python -m apax_b200.build)")
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.apax_abi_version() == 3
    assert handle.apax_gm_workspace_bytes(96, 9000, 1) > 0
    assert handle.apax_nl_workspace_bytes(96, 9000) > 0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "apax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_runtime_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from apax_b200 import runtime

    with pytest.raises(_lib.ApaxB200Error):
        runtime.DeviceParams(syn.gmnn_params())


def test_xla_ffi_shim_is_gated_on_jax():
    """The XLA-FFI layer cannot be compiled without jaxlib's headers; it must say so instead of failing obscurely, and its
    handlers must cover the C entry points they claim to."""
    import os

    from apax_b200.ffi import build_ffi as bf

    src = open(os.path.join(os.path.dirname(bf.__file__), "xla_ffi_shim.cc")).read()
    for fn in ("apax_gm_energy_forces", "apax_gm_descriptor", "apax_gm_descriptor_vjp", "apax_gm_stress_times_vol",
               "apax_gm_stress_times_vol_offsets", "apax_nl_build_minimage", "apax_nl_needs_rebuild", "apax_train_loss_and_grad"):
        assert fn + "(" in src
    # every handler the shim defines is registered by build_ffi.register() (and vice versa)
    import re

    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src))
    assert defined == {sym for _, sym in bf.FFI_TARGETS} and len(defined) == 7
    try:
        import jax  # noqa: F401
    except Exception:
        with pytest.raises(bf.FfiUnavailable):
            bf.build_ffi()


def test_jax_binding_is_importable_and_gated():
    """apax_b200/ffi/jax_binding.py (the jax.ffi.ffi_call / jax.custom_vjp side of the shim) imports without jax, names only
    targets the shim defines, and every entry point raises FfiUnavailable -- not an obscure error -- where jax is absent."""
    import inspect
    import re

    from apax_b200.ffi import jax_binding as jb

    src = inspect.getsource(jb)
    used = set(re.findall(r'ffi_call\("(\w+)"', src))
    assert used == {t for t, _ in jb.FFI_TARGETS}, used ^ {t for t, _ in jb.FFI_TARGETS}
    assert src.count("custom_vjp") >= 4 and "defvjp" in src          # energy and descriptor rules
    try:
        import jax  # noqa: F401
    except Exception:
        for fn, args in ((jb.register, ()), (jb.make_energy_fn, (0, None)), (jb.make_descriptor, (0,)),
                         (jb.neighbor_list_minimage, (None, None, 6.0, 10)), (jb.needs_rebuild, (None, None, None, 0.5))):
            with pytest.raises(jb.FfiUnavailable):
                fn(*args)


def test_public_header_is_plain_c(tmp_path):
    """The drop-in boundary must be bindable from C / cgo / JNI / ctypes: the header compiles as C99 without C++ or CUDA."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "apax_b200.h"\nint main(void) { apax_gm_config c; (void)c; return APAX_B200_ABI_VERSION == 3 ? 0 : 1; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only", str(src)])


def test_builder_empirical_corrections_and_stress_statics():
    """apax/nn/builder.py:146-158: the `empirical_corrections` list of the model config becomes ZBL / exponential entries of
    EnergyModel.corrections; their ranges reach the kernel statics; unknown names are refused (no silent skip)."""
    import pytest

    from apax_b200.nn.builder import GMNNBuilder
    from apax_b200.nn.models import ExponentialRepulsion, ZBLRepulsion

    cfg = {"empirical_corrections": [{"name": "zbl", "r_max": 1.8}, {"name": "exponential", "r_max": 1.5}], "calc_stress": True}
    model = GMNNBuilder(cfg).build_energy_derivative_model(init_box=np.eye(3) * 10.0)
    kinds = [type(c) for c in model.energy_model.corrections]
    assert kinds == [ZBLRepulsion, ExponentialRepulsion]
    st = model.energy_model.statics()
    assert st["zbl_r_max"] == 1.8 and st["exp_r_max"] == 1.5
    assert model.calc_stress
    with pytest.raises(NotImplementedError):
        GMNNBuilder({"empirical_corrections": [{"name": "d3", "r_max": 2.0}]}).build_energy_derivative_model(init_box=np.eye(3))
    plain = GMNNBuilder().build_energy_derivative_model(init_box=np.eye(3) * 10.0)
    assert plain.energy_model.corrections == [] and "zbl_r_max" not in plain.energy_model.statics()


_FFI_STUB = r"""
#pragma once
#include <cstdint>
#include <string>
#include <vector>
typedef struct CUstream_st* cudaStream_t;
namespace xla { namespace ffi {
enum class ErrorCode { kInvalidArgument, kInternal };
struct Error { Error() {} Error(ErrorCode, std::string) {} static Error Success() { return Error(); } };
enum DT { F64, S32, F32, U8 };
template <DT> struct Native;
template <> struct Native<F64> { using t = double; };
template <> struct Native<S32> { using t = int32_t; };
template <> struct Native<F32> { using t = float; };
template <> struct Native<U8> { using t = uint8_t; };
template <DT D> struct Buffer {
  std::vector<int64_t> dimensions() const { return {}; }
  typename Native<D>::t* typed_data() const { return nullptr; }
  size_t element_count() const { return 0; }
};
struct AnyBuffer { size_t element_count() const { return 0; } void* untyped_data() const { return nullptr; } };
template <DT D> struct ResultBuffer { Buffer<D>* operator->() const { return nullptr; } };
struct Traits { static constexpr int kCmdBufferCompatible = 1; };
}}
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(...)
"""


def test_xla_ffi_shim_parses_against_a_stub_of_the_ffi_api(tmp_path):
    """jaxlib's headers are absent, so the shim cannot be built here; a stub of the few FFI types it touches at least keeps its
    handler bodies (argument order and types of every forwarded C-ABI call, include/apax_b200.h) under the compiler's eye."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        import pytest

        pytest.skip("no g++")
    inc = tmp_path / "xla" / "ffi" / "api"
    inc.mkdir(parents=True)
    (inc / "ffi.h").write_text(_FFI_STUB)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", f"-I{tmp_path}", f"-I{os.path.join(root, 'include')}",
                          os.path.join(root, "apax_b200", "ffi", "xla_ffi_shim.cc")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) must print exactly one JSON line on stdout with the
    contract's keys, on a machine without a GPU."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "atom_steps_per_s" and d["unit"] == "atom-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
