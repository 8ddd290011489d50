"""tcgen05.mma kind::i8 issue rate and tensor-memory drain rate, alone and concurrently (SM cycles)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from apax_b200 import runtime as rt

out = torch.zeros(148, 2, device="cuda", dtype=torch.int64)


def run(n, pattern, iters, reads):
    for _ in range(2):
        out.zero_()
        rc = rt.lib().apax_i8_mma_probe(rt._stream_ptr(), n, pattern, iters, reads, rt._ptr(out))
        assert rc == 0
        torch.cuda.synchronize()
    m = out.double().mean(0).tolist()
    return m[0], m[1]


for n in (64, 128, 256):
    for pattern in (0, 1, 2):
        mma, _ = run(n, pattern, 2000, 0)
        cyc = mma / 20000
        print(f"N={n} pattern={pattern}: {cyc:.1f} cycles per MMA  ({128 * n * 32 / cyc:.0f} MAC/clk/SM)")
_, rd = run(128, -1, 0, 200)
print(f"TMEM drain alone: {rd / 200:.0f} cycles per 256 KB pass  ({262144 * 200 / rd:.0f} B/clk/SM)")
for n in (128, 256):
    mma, rd = run(n, 0, 2000, 200)
    print(f"concurrent N={n}: {mma / 20000:.1f} cycles per MMA, drain {rd / 200:.0f} cycles per pass ({262144 * 200 / rd:.0f} B/clk/SM)")
for pattern, label in ((3, "merged schedule, constant data"), (19, "merged schedule, random data"), (16, "N=256 one accumulator, random data")):
    mma, _ = run(256 if pattern == 16 else 64, pattern, 2000, 0)
    per = mma / 2000
    print(f"{label}: {per:.0f} cycles per round" + (" of 4 wide MMAs = 10 digit products (ideal 320)" if pattern != 16 else " of 10 MMAs (ideal 1280)"))
for pattern, label in ((4, "+ commit per 8 MMAs"), (5, "+ tcgen05.fence"), (6, "+ wait on a completed barrier"), (7, "+ wait on the k-block's own commit")):
    mma, _ = run(64, pattern, 2000, 0)
    print(f"merged schedule {label}: {mma / 1000:.0f} cycles per k-block of 8 MMAs (ideal 664)")
