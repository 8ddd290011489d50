"""tcgen05.mma kind::i8 issue-rate probe: SM cycles per MMA for N = 64 / 128 / 256 and three issue patterns."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from apax_b200 import runtime as rt

out = torch.zeros(148, device="cuda", dtype=torch.int64)
for n in (64, 128, 256):
    for pattern in (0, 1, 2):
        for iters in (200, 2000):
            rc = rt.lib().apax_i8_mma_probe(rt._stream_ptr(), n, pattern, iters, rt._ptr(out))
            assert rc == 0
            torch.cuda.synchronize()
        cyc = out.double().mean().item() / (iters * 10)
        print(f"N={n} pattern={pattern}: {cyc:.1f} cycles per MMA  ({128 * n * 32 / cyc:.0f} MAC/clk/SM)")
