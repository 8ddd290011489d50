#!/usr/bin/env python
"""CPU study (no GPU): which stage's fp32 arithmetic carries the force error of the GMNN path?

The CPU oracle is re-run with single stages lifted to fp64 (the per-pair terms, the moment sums, the contractions, the
readout) and the forces compared with the all-fp64 oracle in units of north_star's tolerance (1e-5 |F| + 1e-6 eV/A).  The
result tells which kernels have to carry extra precision for the CUDA path to sit inside the literal tolerance against
the fp64 truth (tests/conftest.py::force_report).  Nothing here is on the product path.

    python tools/study_force_precision.py [--mol 512] [--seed 7]
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

F32, F64 = torch.float32, torch.float64


def staged_energy(params, R, Z, idx, box, st):
    """energy with per-stage dtypes st = dict(cast, pair, mom, contr, readout)"""
    cfg = go.GMNNConfig()
    emb, dense, scale, shift = go._params_to_torch(params)
    offsets = torch.zeros((idx.shape[1], 3), dtype=F64)
    dr_vec = go.compute_distances(R, idx, box, offsets, "minimage").to(st["cast"])   # the reference casts to fp32 here
    dt = st["pair"]
    n_atoms = Z.shape[0]
    i0, i1 = idx[0].long(), idx[1].long()
    Zl = Z.long()
    d = dr_vec.to(dt)
    dr = go.safe_distance(d)
    dn = d / (dr + 1e-5)[:, None]
    c = go.GMNNConfig(descriptor_dtype=dt)
    radial = go.radial_function(dr, Zl[i0.clamp(0, n_atoms - 1)], Zl[i1.clamp(0, n_atoms - 1)], emb.to(dt), c)
    radial = radial * ((idx[0] - idx[1]) != 0).to(dt)[:, None]
    dm = st["mom"]
    m0, m1, m2, m3 = go.geometric_moments(radial.to(dm), dn.to(dm), i1, n_atoms)
    if st.get("mom_round32"):   # the sums carried in fp64 but handed on as fp32 values (what a kernel would store)
        m0, m1, m2, m3 = (m.to(F32) for m in (m0, m1, m2, m3))
    dc = st["contr"]
    m0, m1, m2, m3 = (m.to(dc) for m in (m0, m1, m2, m3))
    c0 = m0
    c1 = torch.einsum("ari,asi->ars", m1, m1)
    c2 = torch.einsum("arij,asij->ars", m2, m2)
    c3 = torch.einsum("arijk,asijk->ars", m3, m3)
    c4 = torch.einsum("arij,asik,atjk->arst", m2, m2, m2)
    c5 = torch.einsum("ari,asj,atij->arst", m1, m1, m2)
    c6 = torch.einsum("arijk,asijl,atkl->arst", m3, m3, m2)
    c7 = torch.einsum("arijk,asij,atk->arst", m3, m2, m1)
    t2, t3 = go.tril_2d_indices(5), go.tril_3d_indices(5)
    g = torch.cat([c0, c1[:, t2[:, 0], t2[:, 1]], c2[:, t2[:, 0], t2[:, 1]], c3[:, t2[:, 0], t2[:, 1]],
                   c4[:, t3[:, 0], t3[:, 1], t3[:, 2]], c5[:, t2[:, 0], t2[:, 1]].reshape(n_atoms, -1),
                   c6[:, t2[:, 0], t2[:, 1]].reshape(n_atoms, -1), c7.reshape(n_atoms, -1)], dim=-1)
    if st.get("g_round32"):
        g = g.to(F32)
    h = go.atomistic_readout(g, dense, go.GMNNConfig(readout_dtype=st["readout"]))
    E_i = go.per_element_scale_shift(h, Z, scale, shift, cfg) * (Z != 0).to(F64)[:, None]
    return torch.sum(E_i.to(F64))


def forces(params, frac, Z, idx, box, st):
    R, Zt, it, bt, _ = go._as_inputs(frac, Z, idx, box, None)
    R = R.clone().requires_grad_(True)
    e = staged_energy(params, R, Zt, it, bt, st)
    (g,) = torch.autograd.grad(e, R)
    return float(e), (-g).numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mol", type=int, default=512)
    ap.add_argument("--seed", type=int, default=7)
    args = ap.parse_args()
    pos, Z, cell = syn.water_box(args.mol, args.seed)
    frac = pos @ np.linalg.inv(cell)
    idx, _, ov = no.jaxmd_sparse_list(frac, cell.T, 6.0, 0.5)
    assert not ov
    params = syn.gmnn_params(seed=1234, random_scale_shift=True, random_bias=True)
    base = dict(cast=F32, pair=F32, mom=F32, contr=F32, readout=F32)
    e64, f64 = forces(params, frac, Z, idx, cell.T, dict(cast=F64, pair=F64, mom=F64, contr=F64, readout=F64))
    tol = 1e-5 * np.abs(f64) + 1e-6
    rows = [("all fp32 (reference arithmetic)", base),
            ("cast fp64 (no fp32 rounding of dr_vec)", dict(base, cast=F64, pair=F64) | dict(pair=F32)),
            ("pair terms fp64", dict(base, cast=F64, pair=F64)),
            ("moment sums fp64", dict(base, mom=F64)),
            ("moment sums fp64, stored fp32", dict(base, mom=F64, mom_round32=True)),
            ("contractions fp64", dict(base, contr=F64)),
            ("readout fp64", dict(base, readout=F64)),
            ("moments+contractions fp64", dict(base, mom=F64, contr=F64)),
            ("moments+contractions fp64, g stored fp32", dict(base, mom=F64, contr=F64, g_round32=True)),
            ("moments+contr+readout fp64", dict(base, mom=F64, contr=F64, readout=F64)),
            ("pair+moments fp64", dict(base, cast=F64, pair=F64, mom=F64)),
            ("all but readout fp64", dict(base, cast=F64, pair=F64, mom=F64, contr=F64)),
            ("all but cast fp64", dict(cast=F32, pair=F64, mom=F64, contr=F64, readout=F64))]
    print(f"{len(Z)} atoms, {idx.shape[1]} list entries, |F| max {np.abs(f64).max():.3f} rms {np.sqrt((f64 ** 2).mean()):.3f}")
    print(f"{'stages in fp64':45s} {'dE/|E|':>10s} {'max |dF|':>10s} {'rms dF':>10s} {'max dF/tol':>11s}")
    for name, st in rows:
        e, f = forces(params, frac, Z, idx, cell.T, st)
        df = f - f64
        print(f"{name:45s} {abs(e - e64) / abs(e64):10.2e} {np.abs(df).max():10.2e} {np.sqrt((df ** 2).mean()):10.2e} "
              f"{(np.abs(df) / tol).max():11.2f}", flush=True)


if __name__ == "__main__":
    main()
