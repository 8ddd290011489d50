"""Time the integer tensor-core GEMM pieces at readout sizes (quantise + GEMM), via the library's per-launch profile."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from apax_b200 import runtime as rt

n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_064
dev = torch.device("cuda")
for k, nc, planes in [(360, 256, 4), (360, 256, 3), (256, 256, 4), (256, 360, 4)]:
    ld = (n_atoms + 127) // 128 * 128
    act = torch.randn(k, ld, device=dev)
    wgt = torch.randn(k, nc, device=dev) / np.sqrt(k)
    out = torch.zeros(nc, ld, device=dev)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    dbg = torch.zeros(148 * 8, device=dev, dtype=torch.int64)

    def go():
        rc = rt.lib().apax_i8_gemm_test(rt._stream_ptr(), rt._ptr(act), k, ld, n_atoms, rt._ptr(wgt), nc, rt._ptr(out), rt._ptr(err), rt._ptr(dbg), planes)
        assert rc == 0

    go(); go()
    rec = rt.profile_kernels(go)
    flops = 2.0 * n_atoms * k * nc
    print(k, nc, [(n, round(ms, 3)) for n, ms in rec], "err", int(err.item()),
          "fp32-equivalent TFLOP/s of GEMM:", round(flops / (rec[-1][1] * 1e-3) / 1e12, 1))
    d = dbg.view(148, 8).double().mean(0).tolist()
    print('   cycles/CTA: mma total %.0f wait_acc %.0f wait_b %.0f wait_a %.0f | epi total %.0f wait %.0f' % tuple(d[:6]))
    del act, out
