import numpy as np, torch, sys
sys.path.insert(0, ".")
from apax_b200 import runtime as rt
dev = torch.device("cuda")
def run(act, wgt, n_atoms, flags, dump=False):
    k, ld = act.shape; nc = wgt.shape[1]
    z = torch.zeros_like
    out = torch.full((nc, ld), -777.0, device=dev); err = torch.zeros(1, device=dev, dtype=torch.int32)
    dbg = torch.full((16386,), -5.0, device=dev)
    rc = rt.lib().apax_tc_gemm_test(rt._stream_ptr(), rt._ptr(act), rt._ptr(z(act)), k, ld, n_atoms, rt._ptr(wgt), rt._ptr(z(wgt)), nc, rt._ptr(out), rt._ptr(err), rt._ptr(dbg) if dump else None, flags)
    torch.cuda.synchronize()
    return rc, int(err.item()), out, dbg
k, nc, ld = 16, 256, 128
act = (torch.arange(k, device=dev).float()[:, None] * 1000 + torch.arange(ld, device=dev).float()[None, :]).contiguous()
wgt = torch.ones(k, nc, device=dev)
rc, err, out, dbg = run(act, wgt, 128, 0, dump=True)
d = dbg.cpu().numpy(); A = d[:2048]
print("rc", rc, "err", err)
for r in range(6): print("A row", r, A[r*32:(r+1)*32].astype(int))
torch.manual_seed(0)
act = torch.randint(1, 4, (k, ld), device=dev).float()
wgt = torch.randint(1, 4, (k, nc), device=dev).float()
ref = wgt.T @ act
for flags in (0, 2, 4, 6):
    rc, err, out, _ = run(act, wgt, 128, flags)
    nz = (out != 0).sum().item()
    match = (out == ref).float().mean().item()
    print(f"flags={flags} rc={rc} err={err} nonzero={nz} exact_match_fraction={match:.4f} sample={out[0,:4].tolist()} ref={ref[0,:4].tolist()}")
