#!/usr/bin/env python
"""CPU study (no GPU): how many base-256 digit planes does the integer tensor-core readout need?

The readout GEMMs (csrc/i8_gemm.cu) write activations and weights as D balanced base-256 digits of a
fixed-point number (|v| < 64 at a per-row power-of-two scale, weights rescaled per input feature so that every weight row
peaks in [32, 64)) and keep the digit products of order s + t <= D - 1, accumulated exactly.  D = 4 is what ships (10 digit
products per MAC).  D = 3 would be 6 products per MAC -- 40 % fewer MMA columns.  This script puts the SAME arithmetic into
the CPU oracle (forward and input-gradient GEMMs of the readout, everything else untouched, accumulation in exact integers
emulated in fp64) and measures energies and forces against the all-fp64 oracle on a 1536-atom periodic water box, next to
the plain fp32 oracle.  Nothing here is on the product path.

    python tools/study_digit_planes.py [--mol 512]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apax_b200 import synthetic as syn  # noqa: E402
from oracle import gmnn_oracle as go  # noqa: E402
from oracle import neighbor_oracle as no  # noqa: E402


def _pow2_scale(m, lo=32.0):
    """power of two s with m / s in [lo, 2 lo) (m > 0); 1 where m == 0"""
    m = np.where(m > 0, m, lo)
    return 2.0 ** np.floor(np.log2(m / lo))


def _digits(a_int, d):
    """balanced base-256 digits, most significant first: a_int = sum_s out[s] * 256**(d-1-s), out[s] in [-128, 127]"""
    out, rest = [], a_int.copy()
    for _ in range(d):
        lo = ((rest + 128) % 256) - 128
        out.append(lo)
        rest = (rest - lo) // 256
    assert np.all(rest == 0), "value does not fit the digits"
    return out[::-1]


def digit_matmul(x, w, d, max_order=None):
    """x [M,K] fp32, w [K,N] fp32 -> fp32 [M,N]: the kernel's arithmetic with d digit planes; digit products of order
    s + t <= max_order are kept (default d - 1, what the kernel does)."""
    max_order = d - 1 if max_order is None else max_order
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    c = _pow2_scale(np.abs(w64).max(axis=1))                    # per input feature: rows of w / c peak in [32, 64)
    wp, xp = w64 / c[:, None], x64 * c[None, :]
    e = _pow2_scale(np.abs(xp).max(axis=1)) * 2.0               # per row: |xp / e| < 32 <= 64
    unit = 2.0 ** (8 * d - 8)
    a_int = np.rint(xp / e[:, None] * unit).astype(np.int64)
    b_int = np.rint(wp * unit).astype(np.int64)
    a_d, b_d = _digits(a_int, d), _digits(b_int, d)
    acc = np.zeros((x.shape[0], w.shape[1]))
    for s in range(d):
        for t in range(min(d, max_order - s + 1)):                # orders s + t <= max_order are kept
            acc += (a_d[s].astype(np.float64) @ b_d[t].astype(np.float64)) * 256.0 ** (2 * d - 2 - s - t)   # exact in fp64 at these sizes
    y = acc / (unit * unit) * e[:, None]
    return y.astype(np.float32)


class DigitMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, w, d, max_order, bwd=None):
        ctx.save_for_backward(w)
        ctx.d, ctx.max_order = bwd or (d, max_order)
        return torch.from_numpy(digit_matmul(h.detach().numpy(), w.detach().numpy(), d, max_order))

    @staticmethod
    def backward(ctx, gy):
        (w,) = ctx.saved_tensors
        gh = digit_matmul(gy.detach().numpy().astype(np.float32), np.ascontiguousarray(w.detach().numpy().T), ctx.d, ctx.max_order)
        return torch.from_numpy(gh), None, None, None, None


def run(d_planes, params, frac, Z, idx, box, max_order=None, bwd=None):
    orig = go.atomistic_readout

    def readout(g, dense, cfg):
        h = g.to(torch.float32)
        for li, layer in enumerate(dense):
            w, b = layer["w"].to(torch.float32), layer["b"].to(torch.float32)
            if li < len(dense) - 1:
                h = go.vp_swish(DigitMM.apply(h, w, d_planes, max_order, bwd) + b)
            else:
                h = h @ w + b                                   # the 256 -> n_out head is an fp32 dot product in the epilogue
        return h

    go.atomistic_readout = readout
    try:
        return go.energy_forces(params, frac, Z, idx, box, None, "minimage")
    finally:
        go.atomistic_readout = orig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mol", type=int, default=512)
    args = ap.parse_args()
    pos, Z, cell = syn.water_box(args.mol, 7)
    frac = pos @ np.linalg.inv(cell)
    idx, _, ov = no.jaxmd_sparse_list(frac, cell.T, 6.0, 0.5)
    assert not ov
    params = syn.gmnn_params(seed=1234, random_scale_shift=True, random_bias=True)
    r64 = go.energy_forces(params, frac, Z, idx, cell.T, None, "minimage", go.fp64_config())
    rows = [("fp32 oracle (reference arithmetic)", go.energy_forces(params, frac, Z, idx, cell.T, None, "minimage"))]
    for d, m in ((4, 3), (4, 2), (3, 2), (3, 3)):
        rows.append((f"{d} digit planes, orders <= {m}", run(d, params, frac, Z, idx, cell.T, m)))
    for bwd in ((3, 3), (4, 2)):   # forward as shipped, only the two input-gradient GEMMs leaner
        rows.append((f"fwd 4/3, bwd {bwd[0]} planes, orders <= {bwd[1]}", run(4, params, frac, Z, idx, cell.T, 3, bwd)))
    tol = 1e-5 * np.abs(r64["forces"]) + 1e-6
    print(f"{len(Z)} atoms, {idx.shape[1]} list entries; errors against the all-fp64 oracle")
    print(f"{'readout arithmetic':40s} {'dE/|E|':>10s} {'max |dF|':>10s} {'rms dF':>10s} {'max dF/tol':>11s}")
    for name, r in rows:
        df = r["forces"] - r64["forces"]
        print(f"{name:40s} {abs(r['energy'] - r64['energy']) / abs(r64['energy']):10.2e} {np.abs(df).max():10.2e} "
              f"{np.sqrt((df ** 2).mean()):10.2e} {(np.abs(df) / tol).max():11.2f}")


if __name__ == "__main__":
    main()
