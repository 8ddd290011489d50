"""apax_gm_descriptor_vjp -- the backward half of the ModelBuilder.build_descriptor seam (apax/nn/builder.py:79-83) --
against torch.autograd through the oracle's restatement of GaussianMomentDescriptor.__call__
(apax/layers/descriptor/gaussian_moment_descriptor.py:31-103), in reference precision and in fp64."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _oracle_vjp(params, dr_vec, Z, idx, g_bar, dtype):
    from oracle import gmnn_oracle as go

    emb = torch.as_tensor(np.asarray(params["params"]["energy_model"]["representation"]["radial_fn"]["atomic_type_embedding"]))
    cfg = go.GMNNConfig(descriptor_dtype=dtype)
    d = torch.as_tensor(np.asarray(dr_vec, dtype=np.float64)).clone().requires_grad_(True)
    g, _ = go.gaussian_moment_descriptor(d, torch.as_tensor(np.asarray(Z).astype(np.int64)), torch.as_tensor(np.asarray(idx).astype(np.int64)),
                                         emb, cfg)
    loss = torch.sum(g.to(torch.float64) * torch.as_tensor(np.asarray(g_bar, dtype=np.float64)))
    (gd,) = torch.autograd.grad(loss, d)
    return g.detach().numpy(), gd.numpy()


def _water_drvec(n_mol, seed):
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no

    pos, Z, cell = syn.water_box(n_mol, seed)
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0, padded_length=int(idx.shape[1] * 1.5))     # (0,0) padding, ase_calc.py:264
    frac = pos @ np.linalg.inv(cell)
    R, Zt, it, bt, ot = go._as_inputs(frac, Z, idx, cell.T, off)
    dr_vec = go.compute_distances(R, it, bt, ot, "offsets").numpy()
    return pos, frac, Z, cell, idx, off, dr_vec


@pytest.mark.parametrize("n_mol,seed", [(32, 0), (11, 3)])
def test_descriptor_vjp_drvec_vs_autograd(n_mol, seed):
    from apax_b200 import runtime as rt
    from apax_b200.layers.descriptor import GaussianMomentDescriptor

    pos, frac, Z, cell, idx, off, dr_vec = _water_drvec(n_mol, seed)
    params = syn.gmnn_params(seed=7)
    emb = params["params"]["energy_model"]["representation"]["radial_fn"]["atomic_type_embedding"]
    g_bar = np.random.default_rng(seed).normal(size=(len(Z), 360)).astype(np.float32)
    desc = GaussianMomentDescriptor()
    variables = {"params": {"radial_fn": {"atomic_type_embedding": emb}}}
    g = desc.apply(variables, dr_vec, Z, idx).cpu().numpy()
    d_bar = desc.vjp(variables, dr_vec, Z, idx, g_bar).cpu().numpy()
    g32, d32 = _oracle_vjp(params, dr_vec, Z, idx, g_bar, torch.float32)
    g64, d64 = _oracle_vjp(params, dr_vec, Z, idx, g_bar, torch.float64)
    assert d_bar.shape == dr_vec.shape and d_bar.dtype == np.float64
    assert np.abs(g - g32).max() <= 2e-6 * np.abs(g32).max()
    scale = np.abs(d64).max()
    err_k, err_o = np.abs(d_bar - d64).max(), np.abs(d32 - d64).max()
    # against reference precision at fp32 rounding level of the largest entry, and no further from fp64 than the oracle is
    assert np.abs(d_bar - d32).max() <= 4e-6 * scale, np.abs(d_bar - d32).max() / scale
    assert err_k <= 1.5 * err_o + 1e-7 * scale, (err_k / scale, err_o / scale)
    pad = idx[0] == idx[1]
    assert pad.any() and np.all(d_bar[pad] == 0.0)                       # padding entries receive exactly nothing


def test_descriptor_vjp_positions_all_branches():
    """R_bar: the cotangent pushed through the displacement (identity Jacobian of space.transform) for the offsets and
    min-image branches, against autograd through the oracle's compute_distances + descriptor; truncated n_contr as well."""
    from apax_b200 import runtime as rt
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no

    pos, frac, Z, cell, idx, off, dr_vec = _water_drvec(24, 5)
    for n_contr in (8, 5):
        n_feat = rt.FEATURE_COUNTS[n_contr - 1]
        params = syn.gmnn_params(seed=3, n_features=n_feat)
        dp = rt.DeviceParams(params, {"n_contr": n_contr})
        g_bar = np.random.default_rng(1).normal(size=(len(Z), n_feat)).astype(np.float32)
        emb = torch.as_tensor(np.asarray(params["params"]["energy_model"]["representation"]["radial_fn"]["atomic_type_embedding"]))
        cases = [("offsets", idx, off)]
        big_pos, big_Z, big_cell = syn.water_box(120, 2)                  # L = 15.3 A > 2 r_max: min-image list is valid
        cases_min = no.jaxmd_sparse_pairs(big_pos @ np.linalg.inv(big_cell), big_cell.T, 6.0)
        for mode, ii, oo in cases + [("minimage", cases_min, None)]:
            fr, zz, cc = (frac, Z, cell) if mode == "offsets" else (big_pos @ np.linalg.inv(big_cell), big_Z, big_cell)
            gb = g_bar if mode == "offsets" else np.random.default_rng(2).normal(size=(len(zz), n_feat)).astype(np.float32)
            d_dr, d_R = rt.descriptor_vjp(dp, fr, zz, ii, gb, cc.T, oo, mode, want_R=True)
            R, Zt, it, bt, ot = go._as_inputs(fr, zz, ii, cc.T, oo)
            outs = {}
            for dt in (torch.float32, torch.float64):
                Rg = R.clone().requires_grad_(True)
                d = go.compute_distances(Rg, it, bt, ot, mode)
                g, _ = go.gaussian_moment_descriptor(d, Zt, it, emb, go.GMNNConfig(descriptor_dtype=dt, n_contr=n_contr))
                (gr,) = torch.autograd.grad(torch.sum(g.to(torch.float64) * torch.as_tensor(gb.astype(np.float64))), Rg)
                outs[dt] = gr.numpy()
            scale = np.abs(outs[torch.float64]).max()
            got = d_R.cpu().numpy()
            err_k, err_o = np.abs(got - outs[torch.float64]).max(), np.abs(outs[torch.float32] - outs[torch.float64]).max()
            assert np.abs(got - outs[torch.float32]).max() <= 4e-6 * scale, (mode, n_contr)
            assert err_k <= 1.5 * err_o + 1e-7 * scale, (mode, n_contr, err_k / scale, err_o / scale)
            assert d_dr.shape == (ii.shape[1], 3)


def test_descriptor_vjp_rejects_bad_arguments():
    from apax_b200 import _lib, runtime as rt

    dp = rt.DeviceParams(syn.gmnn_params(seed=3))
    with pytest.raises(_lib.ApaxB200Error):
        rt.descriptor_vjp(dp, np.zeros((4, 3)), np.ones(3, np.int32), np.zeros((2, 4), np.int32), np.zeros((3, 12), np.float32), mode="drvec")
    with pytest.raises(_lib.ApaxB200Error):   # positions cotangent does not exist when dr_vec itself is the input
        rt.descriptor_vjp(dp, np.zeros((4, 3)), np.ones(3, np.int32), np.zeros((2, 4), np.int32), np.zeros((3, 360), np.float32), mode="drvec",
                          want_R=True)
