"""the Bessel basis (apax's config-file default: n_basis = 16, r_max = 5) and the exponentially
spaced Gaussian basis, evaluated by the fused pair kernels + integer readout (energy+forces) and by the table-driven training kernels
(apax_train_loss_and_grad), against
the fixture written by the reference's OWN BesselBasis / GaussianBasis source (tests/golden/basis_variants.npz) and the
oracle (energy+forces and training gradients)."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn
from conftest import assert_force_parity

pytestmark = pytest.mark.gpu

CASES = {
    "bessel": ({"name": "bessel", "n_basis": 16, "r_max": 5.0}, dict(basis="bessel", n_basis=16, r_max=5.0)),
    "gexp": ({"name": "gaussian", "n_basis": 7, "r_min": 0.5, "r_max": 6.0, "spacing": "exponential"}, dict(spacing="exponential")),
}


@pytest.mark.parametrize("name", ["bessel", "gexp"])
def test_energy_forces_against_reference_source_and_oracle(golden, name):
    from apax_b200.nn.builder import GMNNBuilder
    from oracle import gmnn_oracle as go

    d = golden("basis_variants")
    basis_cfg, ocfg = CASES[name]
    cfg = go.GMNNConfig(**ocfg)
    params = syn.gmnn_params(seed=int(d["param_seed"]), n_basis=cfg.n_basis, random_bias=True, random_scale_shift=True)
    frac, Z, idx, off, cell = d[name + "_frac"], d["Z"], d[name + "_idx"], d[name + "_offsets"], d["cell"]
    model = GMNNBuilder({"basis": basis_cfg}).build_energy_derivative_model(init_box=cell.T)
    out = model.apply(params, frac, Z, idx, cell, off)
    torch.cuda.synchronize()
    e, f = float(out["energy"].item()), out["forces"].cpu().numpy()
    ref64 = go.energy_forces(params, frac, Z, idx, cell, off, "offsets", go.fp64_config(cfg))
    e_ref, f_ref = float(d[name + "_energy"]), d[name + "_forces"]
    # (energies of these random models nearly cancel: accept 1e-6 of |E| or twice the reference's own distance from fp64)
    assert abs(e - ref64["energy"]) <= max(1e-6 * abs(e_ref), 2.0 * abs(e_ref - ref64["energy"]))
    assert_force_parity(f, f_ref, ref64["forces"], where=name)


def test_bessel_training_gradients_vs_oracle():
    from apax_b200.optimizer.get_optimizer import get_opt
    from apax_b200.train.loss import Loss, LossCollection
    from apax_b200.train.trainer import TrainState, TrainStep
    from oracle import gmnn_oracle as go
    from oracle import train_oracle as tor

    cfg = go.GMNNConfig(basis="bessel", n_basis=16, r_max=5.0)
    inputs, labels = tor.ethanol_training_batch(4, seed=6, r_max=5.0, pad_atoms=1, pad_pairs=2)
    params = syn.gmnn_params(seed=8, n_basis=16, random_bias=True, random_scale_shift=True)
    ref = tor.loss_and_grads(params, inputs, labels, cfg=cfg)
    ref64 = tor.loss_and_grads(params, inputs, labels, cfg=go.fp64_config(cfg))
    state = TrainState(params, get_opt(params, 1, 10, emb_lr=0.0, nn_lr=0.0, scale_lr=0.0, shift_lr=0.0),
                       {"n_basis": 16, "r_min": 0.0, "r_max": 5.0, "basis_kind": 2})
    loss_fn = LossCollection([Loss("energy", "mse", 1.0, 1.0), Loss("forces", "mse", 4.0, 1.0)])
    step = TrainStep(state, loss_fn, 4, inputs["positions"].shape[1], inputs["idx"].shape[2])
    loss, pred = step(inputs, labels)
    torch.cuda.synchronize()
    assert abs(float(loss.item()) - ref["loss"]) <= 2e-6 * abs(ref["loss"])
    got = state.grads["params"]["energy_model"]
    flat = {"emb": got["representation"]["radial_fn"]["atomic_type_embedding"], "scale": got["scale_shift"]["scale_per_element"],
            "shift": got["scale_shift"]["shift_per_element"]}
    for i in range(3):
        flat[f"w{i}"], flat[f"b{i}"] = got["readout"][f"dense_{i}"]["w"], got["readout"][f"dense_{i}"]["b"]
    for k, g in flat.items():
        r, t = ref["grads"][k], ref64["grads"][k]
        scale = np.abs(r).max() + 1e-30
        assert np.abs(g - r).max() <= 2e-5 * scale, k
        assert np.abs(g - t).max() <= 2.0 * np.abs(r - t).max() + 2e-6 * scale, k


@pytest.mark.parametrize("name,n_basis", [("bessel", 16), ("bessel", 9), ("gexp", 7), ("glin", 12)])
def test_generic_basis_fast_path_min_image_vs_oracle_and_table_kernels(name, n_basis):
    """The generic-basis pair kernels (coefficients in a [5][16] table, Bessel through one sincos + the Chebyshev recurrence)
    on a 1,536-atom min-image water box: descriptor features, energy and forces against the oracle in reference precision
    and fp64, and against the table-driven kernels (apax_train_energy_forces) that evaluate the basis directly."""
    from apax_b200 import runtime as rt
    from apax_b200.train.trainer import table_model
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no

    ocfg = {"bessel": dict(basis="bessel", n_basis=n_basis, r_max=5.0), "gexp": dict(spacing="exponential", n_basis=n_basis),
            "glin": dict(n_basis=n_basis)}[name]
    cfg = go.GMNNConfig(**ocfg)
    kind = {"bessel": 2, "gexp": 1, "glin": 0}[name]
    pos, Z, cell = syn.water_box(512, 7)
    frac = pos @ np.linalg.inv(cell)
    idx = no.jaxmd_sparse_pairs(frac, cell.T, cfg.r_max)
    params = syn.gmnn_params(seed=17, n_basis=n_basis, random_bias=True, random_scale_shift=True)
    kcfg = {"n_basis": n_basis, "r_min": 0.0 if name == "bessel" else cfg.r_min, "r_max": cfg.r_max, "basis_kind": kind}
    dp = rt.DeviceParams(params, kcfg)
    g = rt.descriptor(dp, frac, Z, idx, cell.T, None, "minimage").cpu().numpy()
    g_ref = go.descriptor_only(params, frac, Z, idx, cell.T, None, "minimage", cfg)["g"]
    assert np.abs(g - g_ref).max() <= 3e-6 * np.abs(g_ref).max(), np.abs(g - g_ref).max() / np.abs(g_ref).max()
    e, f = rt.energy_forces(dp, frac, Z, idx, cell.T, None, "minimage")
    torch.cuda.synchronize()
    e, f = e.cpu().numpy(), f.cpu().numpy()[0]
    ref = go.energy_forces(params, frac, Z, idx, cell.T, None, "minimage", cfg)
    ref64 = go.energy_forces(params, frac, Z, idx, cell.T, None, "minimage", go.fp64_config(cfg))
    assert abs(e[0] - ref64["energy"]) <= max(1e-6 * abs(ref["energy"]), 2.0 * abs(ref["energy"] - ref64["energy"]))
    assert_force_parity(f, ref["forces"], ref64["forces"], where=f"generic basis {name} n_basis {n_basis}")
    # the table-driven kernels on the same model (Cartesian positions, explicit offsets of the min-image list)
    d = frac[idx[1]] - frac[idx[0]]
    offs = (-np.round(d)) @ cell
    e_t, f_t = table_model(params, dict(rt.default_config(), **kcfg)).energy_forces(frac, Z, idx, cell.T, offs, "offsets")
    torch.cuda.synchronize()
    f_t = f_t.cpu().numpy()
    assert abs(float(e_t.cpu().numpy().reshape(-1)[0]) - e[0]) <= 2e-6 * abs(e[0]) + 1e-5
    assert np.abs(f_t - f).max() <= 2e-5 * np.abs(f).max()
