"""The two parity-safe readout implementations against each other: the default integer tensor-core GEMMs
(APAX_B200_READOUT unset / i8) and the fp32 CUDA-core GEMMs (APAX_B200_READOUT=simt, evaluated in a subprocess
because the switch is read once per process).  Both are checked against the oracle elsewhere; here they must
agree with each other at fp32 rounding level on a system large enough to use many atom tiles."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from apax_b200 import runtime as rt, synthetic as syn
n_mol, n_out, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
pos, Z, cell = syn.water_box(n_mol, seed=5)
dp = rt.DeviceParams(syn.gmnn_params(seed=3, n_out=n_out, random_scale_shift=True, random_bias=True), {{"n_out": n_out}})
ctx = rt.HostContext(dp, len(Z), 120 * len(Z))
e, f = ctx.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.0, rebuild=True)
np.savez(out, e=np.asarray(e), f=np.asarray(f))
"""


def _run(mode, n_mol, n_out, tmp_path):
    out = str(tmp_path / f"ef_{mode}.npz")
    env = dict(os.environ)
    if mode:
        env["APAX_B200_READOUT"] = mode
    else:
        env.pop("APAX_B200_READOUT", None)
    subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT), str(n_mol), str(n_out), out], check=True, env=env,
                   timeout=600)
    d = np.load(out)
    return d["e"], d["f"]


@pytest.mark.parametrize("n_mol,n_out", [(1500, 1), (700, 4)])
def test_integer_tensor_core_readout_matches_cuda_core_readout(tmp_path, n_mol, n_out):
    assert torch.cuda.is_available()
    e_i8, f_i8 = _run("", n_mol, n_out, tmp_path)
    e_simt, f_simt = _run("simt", n_mol, n_out, tmp_path)
    assert e_i8.shape == (n_out,) and f_i8.shape == f_simt.shape
    # energies: both are fp64 sums of fp32 per-atom outputs; they differ by the GEMMs' rounding only
    assert np.abs(e_i8 - e_simt).max() <= 2e-7 * np.abs(e_simt).max() + 1e-6
    fmax = np.abs(f_simt).max()
    assert np.abs(f_i8 - f_simt).max() <= 2e-5 * fmax, np.abs(f_i8 - f_simt).max() / fmax
    # and no systematic offset between the two (the truncating 3xTF32 path fails this by two orders of magnitude)
    assert abs(np.mean(f_i8 - f_simt)) <= 1e-7 * fmax
