"""ctypes binding of libapax_b200.so (the C ABI declared in include/apax_b200.h).

There is no CPU fallback: if the library is missing or fails to load, importing this module's
`lib()` raises.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libapax_b200.so")

OK = 0
ERR_INVALID_ARG = -1
ERR_UNSUPPORTED = -2
ERR_WORKSPACE = -3
ERR_CUDA = -4
ERR_OVERFLOW = -5
_ERR_NAMES = {-1: "APAX_ERR_INVALID_ARG", -2: "APAX_ERR_UNSUPPORTED", -3: "APAX_ERR_WORKSPACE", -4: "APAX_ERR_CUDA",
              -5: "APAX_ERR_OVERFLOW"}

DISP_GAS, DISP_OFFSETS, DISP_MINIMAGE, DISP_DRVEC = 0, 1, 2, 3
OPT_READOUT, OPT_PAIR_BLOCK, OPT_MEAN_FORCES, OPT_CONTRACT_SPLIT = 0, 1, 2, 3
READOUT_KINDS = {"simt": 0, "tc": 1, "i8": 2}


class ApaxB200Error(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: {_ERR_NAMES.get(code, code)} {detail}".strip())


class GmConfig(C.Structure):
    _fields_ = [
        ("n_species", C.c_int32), ("n_radial", C.c_int32), ("n_basis", C.c_int32), ("n_contr", C.c_int32),
        ("n_hidden1", C.c_int32), ("n_hidden2", C.c_int32), ("n_out", C.c_int32), ("use_ntk", C.c_int32),
        ("use_embed_norm", C.c_int32), ("mask_atoms", C.c_int32), ("r_min", C.c_float), ("r_max", C.c_float),
        ("basis_kind", C.c_int32),
    ]


class TrainLossConfig(C.Structure):
    _fields_ = [("energy_weight", C.c_double), ("energy_atoms_exponent", C.c_double), ("forces_weight", C.c_double),
                ("forces_atoms_exponent", C.c_double)]


class TrainOptConfig(C.Structure):
    _fields_ = [("emb_lr", C.c_double), ("nn_lr", C.c_double), ("scale_lr", C.c_double), ("shift_lr", C.c_double),
                ("gradient_clipping", C.c_double), ("end_value", C.c_double), ("transition_steps", C.c_int64),
                ("transition_begin", C.c_int64), ("b1", C.c_double), ("b2", C.c_double), ("eps", C.c_double)]


class TrainPeers(C.Structure):
    _fields_ = [("grad32", C.c_void_p * 8), ("grad64", C.c_void_p * 8), ("flags", C.c_void_p * 8), ("world", C.c_int32),
                ("rank", C.c_int32)]


class HaloPeers(C.Structure):
    _fields_ = [("rows", C.c_void_p * 8), ("header", C.c_void_p * 8), ("world", C.c_int32), ("rank", C.c_int32)]


class GmCorrections(C.Structure):
    _fields_ = [("zbl_on", C.c_int32), ("zbl_r_max", C.c_float), ("zbl_coefficients", C.c_float * 4), ("zbl_rep_scale", C.c_float),
                ("exp_on", C.c_int32), ("exp_r_max", C.c_float), ("exp_rep_scale_host", C.c_void_p),
                ("exp_rep_prefactor_host", C.c_void_p)]


class MdNvtConfig(C.Structure):
    _fields_ = [("chain_length", C.c_int32), ("chain_steps", C.c_int32), ("sy_steps", C.c_int32), ("dt", C.c_double),
                ("kT", C.c_double), ("tau", C.c_double)]


# every exported symbol of include/apax_b200.h with its signature
_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
SIGNATURES = {
    "apax_abi_version": (C.c_int, []),
    "apax_last_cuda_error": (C.c_char_p, []),
    "apax_kernel_launch_count": (_I64, []),
    "apax_profile_enable": (C.c_int, [C.c_int]),
    "apax_profile_collect": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "apax_gm_params_create": (C.c_int, [C.POINTER(GmConfig), _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "apax_gm_params_set_corrections": (C.c_int, [_P, C.POINTER(GmCorrections)]),
    "apax_gm_params_set_option": (C.c_int, [_P, _I32, _I32]),
    "apax_gm_params_destroy": (None, [_P]),
    "apax_gm_params_status": (C.c_int, [_P]),
    "apax_gm_workspace_bytes": (C.c_size_t, [_I64, _I64, _I32]),
    "apax_gm_energy_forces": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P, C.c_size_t]),
    "apax_gm_prepare": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _P, C.c_size_t]),
    "apax_gm_prepare_minimage": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, C.c_double, _I64, _P, _P, _P, C.c_size_t, _P, C.c_size_t]),
    "apax_gm_compute": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P, C.c_size_t]),
    "apax_gm_compute_partial": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I32, _P, _P, _P, C.c_size_t]),
    "apax_gm_energy_forces_batched": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P, _P, C.c_size_t]),
    "apax_gm_atomic_energies": (C.c_int, [_P, _P, _I64, _I64, _P, _P, C.c_size_t]),
    "apax_gm_stress_times_vol": (C.c_int, [_P, _P, _I64, _I64, _P, _P, C.c_size_t]),
    "apax_gm_stress_times_vol_offsets": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P, C.c_size_t]),
    "apax_gm_descriptor": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, C.c_size_t]),
    "apax_gm_descriptor_vjp": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P, _P, C.c_size_t]),
    "apax_i8_gemm_test": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _P, _P, _I32]),
    "apax_i8_set_debug": (C.c_int, [_I32, _I32]),
    "apax_i8_mma_probe": (C.c_int, [_P, _I32, _I32, _I32, _I32, _P]),
    "apax_tc_gemm_test": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _P, _I64, _P, _P, _P, _I32]),
    "apax_nl_workspace_bytes": (C.c_size_t, [_I64, _I64]),
    "apax_nl_build_minimage": (C.c_int, [_P, _P, _P, _I64, C.c_double, _I64, _P, _P, _P, _P, C.c_size_t]),
    "apax_nl_build_images": (C.c_int, [_P, _P, _P, _I32, _I64, C.c_double, _I64, _P, _P, _P, _P, _P, _P, C.c_size_t]),
    "apax_nl_batched_workspace_bytes": (C.c_size_t, [_I64, _I64, _I64]),
    "apax_nl_build_images_batched": (C.c_int, [_P, _P, _P, _P, _I64, _I64, C.c_double, _I64, _P, _P, _P, _P, _P, _P, C.c_size_t]),
    "apax_nl_needs_rebuild": (C.c_int, [_P, _P, _P, _P, _I64, C.c_double, _P]),
    "apax_md_nvt_state_bytes": (C.c_size_t, []),
    "apax_md_nvt_init": (C.c_int, [_P, C.POINTER(MdNvtConfig), _P, _P, _I64, _P]),
    "apax_md_nvt_pre_force": (C.c_int, [_P, C.POINTER(MdNvtConfig), _P, _P, _P, _P, _P, _I64, _P]),
    "apax_md_nvt_post_force": (C.c_int, [_P, C.POINTER(MdNvtConfig), _P, _P, _P, _I64, _P]),
    "apax_md_nvt_kick_ke": (C.c_int, [_P, C.POINTER(MdNvtConfig), _P, _P, _P, _I64, _I32, _P, _P]),
    "apax_md_nvt_set_global": (C.c_int, [_P, _P, _P, _I64]),
    "apax_md_nvt_chain_finish": (C.c_int, [_P, C.POINTER(MdNvtConfig), _P, _I64, _P, _P, _I64]),
    "apax_train_plan_create": (C.c_int, [C.POINTER(GmConfig), C.POINTER(_P)]),
    "apax_train_plan_destroy": (None, [_P]),
    "apax_train_param_layout": (C.c_int, [_P, _P, _P]),
    "apax_train_workspace_bytes": (C.c_size_t, [_P, _I64, _I64, _I64]),
    "apax_train_loss_and_grad": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P,
                                           C.POINTER(TrainLossConfig), C.c_double, _P, _P, _P, _P, _P, _P, C.c_size_t]),
    "apax_train_energy_forces": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P, _P, C.c_size_t]),
    "apax_train_adam_step": (C.c_int, [_P, _P, C.POINTER(TrainOptConfig), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "apax_peer_alloc": (C.c_int, [_I64, C.POINTER(_P), _P]),
    "apax_peer_open": (C.c_int, [_P, C.POINTER(_P)]),
    "apax_peer_close": (C.c_int, [_P]),
    "apax_peer_free": (C.c_int, [_P]),
    "apax_train_allreduce_adam_step": (C.c_int, [_P, _P, C.POINTER(TrainOptConfig), C.POINTER(TrainPeers), C.c_uint32, _P, _P, _P, _P,
                                                 _P, _P, _P, _P]),
    "apax_train_peers_status": (C.c_int, [C.POINTER(TrainPeers)]),
    "apax_halo_status": (C.c_int, [C.POINTER(HaloPeers)]),
    "apax_peer_allreduce_sum": (C.c_int, [_P, C.POINTER(HaloPeers), C.c_uint32, _P, _P, _I32]),
    "apax_halo_push_rows": (C.c_int, [_P, _P, _P, _P, C.POINTER(HaloPeers), C.c_uint32]),
    "apax_halo_pull_add_rows": (C.c_int, [_P, _P, _P, _P, C.POINTER(HaloPeers), C.c_uint32, _P, _P, _I32]),
    "apax_ensemble_stats": (C.c_int, [_P, _P, _P, _I64, _I64, _I32, _P, _P, _P, _P, _P]),
    "apax_md_nvt_run": (C.c_int, [_P, C.POINTER(MdNvtConfig), _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, C.c_size_t, _I32]),
    "apax_ctx_create": (C.c_int, [_P, _I64, _I64, C.POINTER(_P)]),
    "apax_ctx_destroy": (None, [_P]),
    "apax_ctx_stress": (C.c_int, [_P, _P]),
    "apax_ctx_calculate": (C.c_int, [_P, _P, _P, _P, _I64, C.c_double, _I32, _P, _P, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ApaxB200Error(ERR_UNSUPPORTED, "load",
                            f"{LIB_PATH} is missing: run `python -m apax_b200.build` (needs nvcc); there is no CPU path")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(code: int, where: str) -> None:
    if code != OK:
        detail = ""
        if code == ERR_CUDA:
            detail = lib().apax_last_cuda_error().decode()
        raise ApaxB200Error(code, where, detail)
