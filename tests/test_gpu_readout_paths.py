"""The two parity-safe readout implementations against each other: the default integer tensor-core GEMMs and the fp32
CUDA-core GEMMs (config {"readout": "simt"} -> apax_gm_params_set_option(APAX_OPT_READOUT); the switch is a field of the
parameter handle, nothing is read from the environment).  Both are checked against the oracle elsewhere; here they must
agree with each other at fp32 rounding level on a system large enough to use many atom tiles."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _run(mode, n_mol, n_out):
    from apax_b200 import runtime as rt

    pos, Z, cell = syn.water_box(n_mol, seed=5)
    cfg = {"n_out": n_out}
    if mode:
        cfg["readout"] = mode
    dp = rt.DeviceParams(syn.gmnn_params(seed=3, n_out=n_out, random_scale_shift=True, random_bias=True), cfg)
    ctx = rt.HostContext(dp, len(Z), 120 * len(Z))
    e, f = ctx.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.0, rebuild=True)
    ctx.close()
    return np.asarray(e), np.asarray(f)


@pytest.mark.parametrize("n_mol,n_out", [(1500, 1), (700, 4)])
def test_integer_tensor_core_readout_matches_cuda_core_readout(n_mol, n_out):
    assert torch.cuda.is_available()
    e_i8, f_i8 = _run("", n_mol, n_out)
    e_simt, f_simt = _run("simt", n_mol, n_out)
    assert e_i8.shape == (n_out,) and f_i8.shape == f_simt.shape
    # energies: both are fp64 sums of fp32 per-atom outputs; they differ by the GEMMs' rounding only
    assert np.abs(e_i8 - e_simt).max() <= 2e-7 * np.abs(e_simt).max() + 1e-6
    fmax = np.abs(f_simt).max()
    assert np.abs(f_i8 - f_simt).max() <= 2e-5 * fmax, np.abs(f_i8 - f_simt).max() / fmax
    # and no systematic offset between the two (the truncating 3xTF32 path fails this by two orders of magnitude)
    assert abs(np.mean(f_i8 - f_simt)) <= 1e-7 * fmax


def test_option_rejects_unknown_values():
    from apax_b200 import _lib, runtime as rt

    dp = rt.DeviceParams(syn.gmnn_params(seed=3))
    assert rt.lib().apax_gm_params_set_option(dp.handle, _lib.OPT_READOUT, 7) == _lib.ERR_INVALID_ARG
    assert rt.lib().apax_gm_params_set_option(dp.handle, 99, 0) == _lib.ERR_INVALID_ARG
    assert rt.lib().apax_gm_params_set_option(dp.handle, _lib.OPT_PAIR_BLOCK, 0) == _lib.OK


@pytest.mark.parametrize("n_mol,n_out", [(1500, 1), (45, 1), (300, 3)])
def test_split_contraction_kernels_match_thread_per_atom_kernels(n_mol, n_out):
    """Small systems run the contraction kernels with one atom on four warps (APAX_OPT_CONTRACT_SPLIT, default up to 8,192
    central atoms); config {"contract_split": 0} forces the thread-per-atom kernels.  Every feature is computed by the same
    instruction sequence in both forms, so the digit planes and the energies are bitwise identical; the moment adjoints
    are summed in a different order, so forces agree at fp32 rounding level."""
    from apax_b200 import runtime as rt

    pos, Z, cell = syn.water_box(n_mol, seed=7)
    out = {}
    for split in (0, 1 << 20):
        dp = rt.DeviceParams(syn.gmnn_params(seed=3, n_out=n_out, random_scale_shift=True, random_bias=True),
                             {"n_out": n_out, "contract_split": split})
        ctx = rt.HostContext(dp, len(Z), 120 * len(Z))
        e, f = ctx.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.0, rebuild=True)
        ctx.close()
        out[split] = (np.asarray(e).copy(), np.asarray(f).copy())
    (e0, f0), (e1, f1) = out[0], out[1 << 20]
    assert np.array_equal(e0, e1)
    fmax = np.abs(f0).max()
    assert np.isfinite(f1).all() and np.abs(f1 - f0).max() <= 2e-6 * fmax, np.abs(f1 - f0).max() / fmax
