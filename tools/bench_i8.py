"""Time the integer tensor-core GEMM pieces at readout sizes (quantise + GEMM), via the library's per-launch profile."""
import sys
import numpy as np
import torch
from apax_b200 import runtime as rt

n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_064
dev = torch.device("cuda")
for k, nc in [(360, 256), (256, 256), (256, 360)]:
    ld = (n_atoms + 127) // 128 * 128
    act = torch.randn(k, ld, device=dev)
    wgt = torch.randn(k, nc, device=dev) / np.sqrt(k)
    out = torch.zeros(nc, ld, device=dev)
    err = torch.zeros(1, device=dev, dtype=torch.int32)

    def go():
        rc = rt.lib().apax_i8_gemm_test(rt._stream_ptr(), rt._ptr(act), k, ld, n_atoms, rt._ptr(wgt), nc, rt._ptr(out), rt._ptr(err))
        assert rc == 0

    go(); go()
    rec = rt.profile_kernels(go)
    flops = 2.0 * n_atoms * k * nc
    print(k, nc, [(n, round(ms, 3)) for n, ms in rec], "err", int(err.item()),
          "fp32-equivalent TFLOP/s of GEMM:", round(flops / (rec[-1][1] * 1e-3) / 1e12, 1))
    del act, out
