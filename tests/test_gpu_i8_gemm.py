"""The exact integer tensor-core readout GEMM (tcgen05.mma kind::i8, digit planes) against an fp64 matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(act, wgt, n_atoms, planes=4):
    from apax_b200 import runtime as rt

    k, ld = act.shape
    nc = wgt.shape[1]
    out = torch.zeros(nc, ld, device=act.device, dtype=torch.float32)
    err = torch.zeros(1, device=act.device, dtype=torch.int32)
    rc = rt.lib().apax_i8_gemm_test(rt._stream_ptr(), rt._ptr(act), k, ld, n_atoms, rt._ptr(wgt), nc, rt._ptr(out), rt._ptr(err), None, planes)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "tensor-core pipeline timed out"
    return out


@pytest.mark.parametrize("k,nc,n_atoms,planes", [(32, 128, 128, 4), (64, 64, 128, 4), (256, 256, 1000, 4), (360, 256, 4096, 3),
                                                 (360, 256, 4096, 4), (256, 360, 777, 4), (256, 256, 40000, 4)])
def test_i8_gemm_matches_fp64(k, nc, n_atoms, planes):
    torch.manual_seed(0)
    dev = torch.device("cuda")
    ld = (n_atoms + 127) // 128 * 128
    act = torch.randn(k, ld, device=dev, dtype=torch.float32)
    # feature magnitudes spread over six decades, compensated by the weights (like moments of different rank)
    spread = 10.0 ** torch.linspace(-3, 3, k, device=dev)[:, None]
    act = act * spread
    wgt = torch.randn(k, nc, device=dev, dtype=torch.float32) / np.sqrt(k) / spread
    out = _run(act, wgt, n_atoms, planes)
    ref = (wgt.double().T @ act.double())[:, :n_atoms]
    got = out[:, :n_atoms].double()
    scale = (wgt.double().abs().T @ act.double().abs())[:, :n_atoms]
    rel = ((got - ref).abs() / scale).max().item()
    # fp32 result of the same product for comparison
    f32 = (wgt.T @ act)[:, :n_atoms].double()
    rel32 = ((f32 - ref).abs() / scale).max().item()
    bias = ((got - ref).sum() / scale.sum()).item()
    print(f"k={k} nc={nc} planes={planes}: max err / sum|a||b| = {rel:.2e} (torch fp32 matmul {rel32:.2e}), mean signed {bias:.2e}")
    # four planes: final fp32 rounding of the output only (2^-24 = 6e-8 relative to the result itself);
    # three planes: activations carry 24 bits below the row maximum, like an fp32 significand
    assert rel < (6e-8 if planes == 4 else 3e-7)
    assert abs(bias) < 3e-9    # and no systematic bias, unlike the truncating fp32 tensor-core accumulator
    assert torch.all(out[:, n_atoms:] == 0)   # padding atoms quantise to zero rows


def test_i8_gemm_zero_and_tiny_rows():
    torch.manual_seed(1)
    dev = torch.device("cuda")
    k, nc, n_atoms = 128, 128, 256
    act = torch.randn(k, n_atoms, device=dev, dtype=torch.float32)
    act[:, 3] = 0.0
    act[:, 5] *= 1e-30
    act[:, 7] *= 1e30
    wgt = torch.randn(k, nc, device=dev, dtype=torch.float32)
    wgt[11, :] = 0.0
    out = _run(act, wgt, n_atoms).double()
    ref = wgt.double().T @ act.double()
    scale = wgt.double().abs().T @ act.double().abs()
    assert torch.all(out[:, 3] == 0)
    ok = scale > 0
    assert ((out - ref).abs()[ok] / scale[ok]).max().item() < 1e-7
