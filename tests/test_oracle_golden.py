"""The oracle against fixtures produced by executing the reference's OWN source files
(tests/golden/make_golden.py, under the torch-backed jax/flax shim).  CPU only."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn
from oracle import gmnn_oracle as go
from oracle import neighbor_oracle as no


def test_descriptor_unit_case_is_bit_exact(golden):
    d = golden("descriptor_unit")
    p = syn.gmnn_params(seed=int(d["param_seed"]))
    emb = torch.as_tensor(p["params"]["energy_model"]["representation"]["radial_fn"]["atomic_type_embedding"])
    g, _ = go.gaussian_moment_descriptor(torch.as_tensor(d["dR"]), torch.as_tensor(d["Z"]), torch.as_tensor(d["idx"]), emb,
                                         go.GMNNConfig())
    assert g.shape == (3, 360)  # tests/unit_tests/layers/descriptor/test_gaussian_moment_descriptor.py:30
    assert np.array_equal(g.numpy(), d["g"])


def test_gas_phase_padded_and_unpadded(golden):
    d = golden("gas_padded")
    p = syn.gmnn_params(seed=11, scale=d["scale"], shift=d["shift"])
    r = go.energy_forces(p, d["R"], d["Z"], d["idx"], np.zeros((3, 3)), None, "gas",
                         go.GMNNConfig(apply_mask=False, mask_atoms=False))
    assert abs(r["energy"] - d["energy"]) < 1e-6 * abs(d["energy"])
    assert np.abs(r["forces"] - d["forces"]).max() < 1e-6
    rp = go.energy_forces(p, d["Rp"], d["Zp"], d["idxp"], np.zeros((3, 3)), None, "gas")
    assert abs(rp["energy"] - d["energy_padded"]) < 1e-6 * abs(d["energy"])
    assert np.abs(rp["forces"] - d["forces_padded"]).max() < 1e-6
    # the reference's own assertion (tests/unit_tests/model/test_apax.py:61-62)
    assert abs(r["energy"] - rp["energy"]) < 1e-6
    assert np.all(np.abs(r["forces"] - rp["forces"][:-1]) < 1e-6)


def test_multi_image_water(golden):
    d = golden("water_multi_image")
    p = syn.gmnn_params(seed=int(d["param_seed"]))
    r = go.energy_forces(p, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"], "offsets")
    assert abs(r["energy"] - d["energy"]) < 1e-6 * abs(d["energy"])
    assert np.all(np.abs(r["forces"] - d["forces"]) <= 1e-5 * np.abs(d["forces"]) + 1e-6)


def test_min_image_water_and_neighbor_list(golden):
    d = golden("water_min_image")
    p = syn.gmnn_params(seed=int(d["param_seed"]))
    r = go.energy_forces(p, d["frac"], d["Z"], d["idx"], d["box"], None, "minimage")
    assert abs(r["energy"] - d["energy"]) < 1e-6 * abs(d["energy"])
    assert np.all(np.abs(r["forces"] - d["forces"]) <= 1e-5 * np.abs(d["forces"]) + 1e-6)
    # neighbour list: bit-exact indices, order and padding against partition.py run under the shim
    idx, occ, ov = no.jaxmd_sparse_list(d["frac"], d["box"], 6.0, float(d["skin"]))
    assert np.array_equal(idx, d["idx"])
    assert occ == int(d["n_valid"]) and int(ov) == int(d["overflow"])
    assert no.needs_rebuild(d["moved_small"], d["frac"], d["box"], float(d["skin"])) == bool(d["rebuilt_small"])
    assert no.needs_rebuild(d["moved_big"], d["frac"], d["box"], float(d["skin"])) == bool(d["rebuilt_big"])
    idxb, _, _ = no.jaxmd_sparse_list(d["moved_big"], d["box"], 6.0, float(d["skin"]), max_occupancy=d["idx"].shape[1])
    assert np.array_equal(idxb, d["idx_big"])


def test_shallow_ensemble(golden):
    d = golden("water_ensemble")
    p = syn.gmnn_params(seed=int(d["param_seed"]), n_out=int(d["n_ens"]), random_scale_shift=True, random_bias=True)
    r = go.shallow_ensemble(p, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"], "offsets")
    for k in ("energy", "energy_ensemble", "energy_uncertainty"):
        assert np.allclose(r[k], d[k], rtol=2e-6, atol=1e-6), k
    for k in ("forces", "forces_uncertainty", "forces_ensemble"):
        assert np.all(np.abs(r[k] - d[k]) <= 1e-5 * np.abs(d[k]) + 2e-6), k


def test_md_integrator_against_reference_source(golden):
    """oracle/md_oracle.py vs jax_md's nvt_nose_hoover executed from the reference's own simulate.py
    (tests/golden/make_golden.py::case_md_nvt): 7 steps, triclinic box, unequal masses."""
    from oracle import md_oracle as mo

    d = golden("md_nvt")
    box = d["cell"].T.copy()
    anchor, k = d["frac0"].copy(), float(d["kspring"])

    def force(R):
        dd = R - anchor
        dd = dd - np.round(dd)
        return -k * (dd @ box.T)

    st = mo.nvt_init(d["frac0"], d["p0"], d["mass"], force, float(d["kT"]), float(d["dt"]))
    assert st.chain.KE == float(d["KE0"])
    for step in range(1, 8):
        st = mo.nvt_step(st, force, box, float(d["kT"]), float(d["dt"]))
        if step in (1, 7):
            assert np.abs(st.position - d[f"position_{step}"]).max() < 1e-14
            assert np.abs(st.momentum - d[f"momentum_{step}"]).max() < 1e-13
            assert np.abs(st.chain.xi - d[f"xi_{step}"]).max() < 1e-15
            assert np.abs(st.chain.p_xi - d[f"p_xi_{step}"]).max() < 1e-13
            assert np.array_equal(st.chain.Q, d[f"Q_{step}"])
            assert abs(st.chain.KE - float(d[f"KE_{step}"])) < 1e-12


@pytest.mark.parametrize("name,kw", [("bessel", dict(basis="bessel", n_basis=16, r_max=5.0)), ("gexp", dict(spacing="exponential"))])
def test_basis_variants_against_reference_source(golden, name, kw):
    """BesselBasis (basis_functions.py:59-79) and GaussianBasis(spacing="exponential") (:29-37): raw basis values bit-exact,
    E+F to fp32 rounding, against the fixture written by the reference's own classes."""
    import torch

    from apax_b200 import synthetic as syn
    from oracle import gmnn_oracle as go

    d = golden("basis_variants")
    cfg = go.GMNNConfig(**kw)
    fn = go.bessel_basis if name == "bessel" else go.gaussian_basis
    assert np.array_equal(fn(torch.as_tensor(d["dr_grid"]), cfg).numpy(), d[name + "_basis"])
    params = syn.gmnn_params(seed=int(d["param_seed"]), n_basis=cfg.n_basis, random_bias=True, random_scale_shift=True)
    ref = go.energy_forces(params, d[name + "_frac"], d["Z"], d[name + "_idx"], d["cell"], d[name + "_offsets"], "offsets", cfg)
    assert np.abs(ref["forces"] - d[name + "_forces"]).max() <= 1e-6 * np.abs(d[name + "_forces"]).max() + 1e-7
    assert abs(ref["energy"] - float(d[name + "_energy"])) <= 1e-5


def test_stress_oracle_against_reference_source(golden):
    """stress_times_vol (apax/layers/properties.py:11-42) restated in the oracle vs the fixture from the reference's own source."""
    from apax_b200 import synthetic as syn
    from oracle import gmnn_oracle as go

    d = golden("water_stress")
    params = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True)
    st = go.stress_times_vol(params, d["frac"], d["Z"], d["idx"], d["box"], None, "minimage")
    assert np.abs(st - d["stress"]).max() <= 1e-6 * np.abs(d["stress"]).max()
    # symmetric up to rounding (rotational invariance), and the offsets branch's doubled displacement is a different quantity
    assert np.abs(st - st.T).max() <= 1e-5 * np.abs(st).max()


def test_stress_offsets_oracle_against_reference_source(golden):
    """The offsets branch of stress_times_vol as the reference's disp_fn computes it (apax/layers/distances.py:12-22: the strain
    acts on dR + (I + eps) . dR, offsets added afterwards) vs the fixture from the reference's own source."""
    from apax_b200 import synthetic as syn
    from oracle import gmnn_oracle as go

    d = golden("water_stress_offsets")
    params = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True)
    st = go.stress_times_vol(params, d["frac"], d["Z"], d["idx"], d["cell"], d["offsets"], "offsets")
    assert np.abs(st - d["stress"]).max() <= 1e-6 * np.abs(d["stress"]).max()
    assert np.abs(st - st.T).max() > 1e-2 * np.abs(st).max()      # not a physical stress: the doubled displacement breaks the symmetry


def test_empirical_corrections_against_reference_source(golden):
    """ZBLRepulsion / ExponentialRepulsion (apax/layers/empirical.py:25-126) restated in the oracle vs the reference's own source."""
    from apax_b200 import synthetic as syn
    from oracle import gmnn_oracle as go

    d = golden("water_corrections")
    params = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True, random_scale_shift=True)
    em = params["params"]["energy_model"]
    em["corrections_0"] = {"coefficients": d["zbl_coefficients"], "rep_scale": d["zbl_rep_scale"]}
    em["corrections_1"] = {"rep_scale": d["exp_rep_scale"], "rep_prefactor": d["exp_rep_prefactor"]}
    cfg = go.GMNNConfig(zbl_r_max=float(d["zbl_r_max"]), exp_r_max=float(d["exp_r_max"]))
    ref = go.energy_forces(params, d["frac"], d["Z"], d["idx"], d["cell"], d["offsets"], "offsets", cfg)
    assert abs(ref["energy"] - float(d["energy"])) <= 1e-7 * abs(float(d["energy"]))
    assert np.abs(ref["forces"] - d["forces"]).max() <= 1e-6 * np.abs(d["forces"]).max()


def _water_with_corrections(n_mol=8, seed=3):
    from apax_b200 import synthetic as syn
    from oracle import gmnn_oracle as go, neighbor_oracle as no

    pos, Z, cell = syn.water_box(n_mol, seed)
    idx, offsets = no.vesin_idx_offsets(pos, cell, 6.0)
    frac = pos @ np.linalg.inv(cell)
    rng = np.random.default_rng(seed)
    p = syn.gmnn_params(seed=17, random_bias=True, random_scale_shift=True)
    em = p["params"]["energy_model"]
    em["corrections_0"] = {"coefficients": rng.normal(size=(4, 1)).astype(np.float32), "rep_scale": np.array([-1.2], dtype=np.float32)}
    em["corrections_1"] = {"rep_scale": (rng.uniform(0.3, 1.2, size=119) * rng.choice([-1, 1], size=119)).astype(np.float32),
                           "rep_prefactor": rng.uniform(1.0, 20.0, size=119).astype(np.float32)}
    cfg = go.fp64_config(go.GMNNConfig(zbl_r_max=1.8, exp_r_max=1.6))
    cfg.correction_dtype = torch.float64      # finite differences need the whole energy in fp64 (the reference's terms run in fp32)
    return go, p, frac, Z, idx, cell, offsets, cfg


def test_oracle_corrections_forces_match_finite_differences():
    """Self-check of the oracle's correction terms: fp64 forces vs central differences of the fp64 energy, and the
    terms really act (the energy without the correction blocks differs)."""
    go, p, frac, Z, idx, cell, offsets, cfg = _water_with_corrections()
    ref = go.energy_forces(p, frac, Z, idx, cell, offsets, "offsets", cfg)
    h = 1e-6
    rng = np.random.default_rng(0)
    for a, k in zip(rng.integers(0, len(Z), 6), rng.integers(0, 3, 6)):
        dpos = np.zeros_like(frac)
        dpos[a, k] = h
        dfrac = dpos @ np.linalg.inv(cell)           # a Cartesian step of h along k, in fractional coordinates
        ep = go.energy_forces(p, frac + dfrac, Z, idx, cell, offsets, "offsets", cfg)["energy"]
        em_ = go.energy_forces(p, frac - dfrac, Z, idx, cell, offsets, "offsets", cfg)["energy"]
        fd = -(ep - em_) / (2 * h)
        assert abs(fd - ref["forces"][a, k]) <= 1e-5 * max(1.0, abs(fd)), (a, k, fd, ref["forces"][a, k])
    plain = {"params": {"energy_model": {k: v for k, v in p["params"]["energy_model"].items() if not k.startswith("corrections_")}}}
    assert abs(go.energy_forces(plain, frac, Z, idx, cell, offsets, "offsets", cfg)["energy"] - ref["energy"]) > 1.0


def test_oracle_stress_with_corrections_matches_strain_finite_differences():
    """stress_times_vol of the oracle on the min-image branch incl. the correction terms vs central differences in the strain
    (dr' = (I + eps) . dr applied to the box: box' = (I + eps) . box leaves the fractional coordinates alone)."""
    from oracle import neighbor_oracle as no

    go, p, frac, Z, _idx, cell, _off, cfg = _water_with_corrections(n_mol=64, seed=4)
    idx, occ, ov = no.jaxmd_sparse_list(frac, cell.T, 6.0, 0.0)
    assert not ov
    st = go.stress_times_vol(p, frac, Z, idx, cell.T, None, "minimage", cfg)
    h = 1e-6
    for (i, j) in ((0, 0), (0, 1), (2, 1)):
        eps = np.zeros((3, 3))
        eps[i, j] = h
        ep = go.energy_forces(p, frac, Z, idx, (np.eye(3) + eps) @ cell.T, None, "minimage", cfg)["energy"]
        em_ = go.energy_forces(p, frac, Z, idx, (np.eye(3) - eps) @ cell.T, None, "minimage", cfg)["energy"]
        fd = (ep - em_) / (2 * h)
        assert abs(fd - st[i, j]) <= 1e-5 * max(1.0, np.abs(st).max()), (i, j, fd, st[i, j])
