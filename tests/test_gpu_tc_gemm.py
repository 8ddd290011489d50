"""The tcgen05 3xTF32 readout GEMM against an fp64 matmul."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rna(x: torch.Tensor):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _split(x: torch.Tensor):
    hi = _rna(x)
    return hi, _rna(x - hi)


@pytest.mark.parametrize("k,nc,n_atoms", [(16, 256, 128), (256, 256, 1000), (360, 256, 4096), (256, 384, 777)])
def test_tc_gemm_matches_fp64(k, nc, n_atoms):
    from apax_b200 import runtime as rt

    torch.manual_seed(0)
    dev = torch.device("cuda")
    ld = (n_atoms + 127) // 128 * 128
    act = torch.randn(k, ld, device=dev, dtype=torch.float32)
    wgt = torch.randn(k, nc, device=dev, dtype=torch.float32) / np.sqrt(k)
    a_hi, a_lo = _split(act)
    w_hi, w_lo = _split(wgt)
    assert (a_hi + a_lo - act).abs().max() <= 2.0 ** -21 * act.abs().max()
    out = torch.zeros(nc, ld, device=dev, dtype=torch.float32)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    rc = rt.lib().apax_tc_gemm_test(rt._stream_ptr(), rt._ptr(a_hi), rt._ptr(a_lo), k, ld, n_atoms, rt._ptr(w_hi), rt._ptr(w_lo),
                                    nc, rt._ptr(out), rt._ptr(err), None, 0)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "tensor-core pipeline timed out"
    ref = (wgt.double().T @ act.double())          # [nc][ld]
    n_tiles = (n_atoms + 127) // 128 * 128
    got = out[:, :n_tiles].double()
    scale = (wgt.double().abs().T @ act.double().abs())[:, :n_tiles]
    rel = ((got - ref[:, :n_tiles]).abs() / scale).max().item()
    print(f"k={k} nc={nc}: max error relative to sum|a||b| = {rel:.2e}")
    # single-pass tf32 would be ~5e-4; 3xTF32 removes the operand rounding, what remains is the tensor core's
    # truncating fp32 accumulator (error grows with the number of accumulation steps, k/8 * 3)
    assert rel < 2e-6
