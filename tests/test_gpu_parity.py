"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances are north_star's: neighbour indices bit-exact, fp32 energies within 1e-6 relative,
forces within 1e-5 relative + 1e-6 eV/A absolute."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn
from conftest import assert_energy_parity, assert_force_parity, force_tolerance_ok, max_force_violation

pytestmark = pytest.mark.gpu

E_RTOL = 1e-6


@pytest.fixture(scope="module")
def rt():
    from apax_b200 import runtime

    assert torch.cuda.is_available()
    return runtime


@pytest.fixture(scope="module")
def oracle():
    from oracle import gmnn_oracle, neighbor_oracle

    return gmnn_oracle, neighbor_oracle


def _ef(rt, params, R, Z, idx, box, offsets, mode, cfg=None):
    dp = rt.DeviceParams(params, cfg)
    e, f = rt.energy_forces(dp, R, Z, idx, box, offsets, mode)
    torch.cuda.synchronize()
    return e.cpu().numpy(), f.cpu().numpy()


def test_descriptor_unit_case(rt, golden):
    d = golden("descriptor_unit")
    dp = rt.DeviceParams(syn.gmnn_params(seed=int(d["param_seed"])))
    g = rt.descriptor(dp, d["dR"].astype(np.float64), d["Z"], d["idx"], mode="drvec").cpu().numpy()
    assert g.shape == (3, 360)
    assert np.abs(g - d["g"]).max() <= 2e-6 * np.abs(d["g"]).max()


def test_gas_phase_golden_padding_invariance(rt, golden):
    d = golden("gas_padded")
    p = syn.gmnn_params(seed=11, scale=d["scale"], shift=d["shift"])
    e, f = _ef(rt, p, d["R"], d["Z"], d["idx"], None, None, "gas", {"mask_atoms": False})
    assert abs(e[0] - d["energy"]) <= E_RTOL * abs(d["energy"])
    assert force_tolerance_ok(f[0], d["forces"]), max_force_violation(f[0], d["forces"])
    ep, fp = _ef(rt, p, d["Rp"], d["Zp"], d["idxp"], None, None, "gas")
    assert abs(ep[0] - d["energy_padded"]) <= E_RTOL * abs(d["energy_padded"])
    assert force_tolerance_ok(fp[0], d["forces_padded"])
    # tests/unit_tests/model/test_apax.py:61-62
    assert abs(e[0] - ep[0]) < 1e-6 and np.all(np.abs(f[0] - fp[0][:-1]) < 1e-6)
    assert np.all(fp[0][-1] == 0.0)  # the padded atom feels exactly nothing


def test_multi_image_golden(rt, golden, oracle):
    go, _ = oracle
    d = golden("water_multi_image")
    p = syn.gmnn_params(seed=int(d["param_seed"]))
    e, f = _ef(rt, p, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"], "offsets")
    assert abs(e[0] - d["energy"]) <= E_RTOL * abs(d["energy"]), (e[0], d["energy"])
    ref64 = go.energy_forces(p, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"], "offsets", go.fp64_config())
    assert_force_parity(f[0], d["forces"], ref64["forces"], where="golden multi-image")


def test_min_image_golden(rt, golden, oracle):
    go, _ = oracle
    d = golden("water_min_image")
    p = syn.gmnn_params(seed=int(d["param_seed"]))
    e, f = _ef(rt, p, d["frac"], d["Z"], d["idx"], d["box"], None, "minimage")
    assert abs(e[0] - d["energy"]) <= E_RTOL * abs(d["energy"]), (e[0], d["energy"])
    ref64 = go.energy_forces(p, d["frac"], d["Z"], d["idx"], d["box"], None, "minimage", go.fp64_config())
    assert_force_parity(f[0], d["forces"], ref64["forces"], where="golden min-image")


def test_shallow_ensemble_golden(rt, golden, oracle):
    go, _ = oracle
    d = golden("water_ensemble")
    n_ens = int(d["n_ens"])
    p = syn.gmnn_params(seed=int(d["param_seed"]), n_out=n_ens, random_scale_shift=True, random_bias=True)
    e, f = _ef(rt, p, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"], "offsets")
    f_ens = np.transpose(f, (1, 2, 0))
    r64 = go.shallow_ensemble(p, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"], "offsets", go.fp64_config())
    assert_energy_parity(e, d["energy_ensemble"], r64["energy_ensemble"], where="golden ensemble")
    assert_force_parity(f_ens, d["forces_ensemble"], r64["forces_ensemble"], where="golden ensemble")


@pytest.mark.parametrize("n_mol,seed", [(32, 0), (32, 1), (10, 4)])
def test_water_multi_image_vs_oracle(rt, oracle, n_mol, seed):
    go, no = oracle
    pos, Z, cell = syn.water_box(n_mol, seed)
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0, padded_length=int(idx.shape[1] * 1.5))  # ase_calc.py:264
    frac = pos @ np.linalg.inv(cell)
    p = syn.gmnn_params(seed=1234)
    ref = go.energy_forces(p, frac, Z, idx, cell.T, off, "offsets")
    e, f = _ef(rt, p, frac, Z, idx, cell.T, off, "offsets")
    assert abs(e[0] - ref["energy"]) <= E_RTOL * abs(ref["energy"]), (e[0], ref["energy"])
    ref64 = go.energy_forces(p, frac, Z, idx, cell.T, off, "offsets", go.fp64_config())
    assert_force_parity(f[0], ref["forces"], ref64["forces"], where=f"water {3 * n_mol}")


def test_water_1536_min_image_vs_oracle(rt, oracle):
    go, no = oracle
    pos, Z, cell = syn.water_box(512, 7)
    frac = pos @ np.linalg.inv(cell)
    idx, occ, ov = no.jaxmd_sparse_list(frac, cell.T, 6.0, 0.5)
    assert not ov
    for p in (syn.gmnn_params(seed=1234, random_scale_shift=True, random_bias=True), syn.gmnn_params(seed=1234)):
        ref = go.energy_forces(p, frac, Z, idx, cell.T, None, "minimage")
        ref64 = go.energy_forces(p, frac, Z, idx, cell.T, None, "minimage", go.fp64_config())
        e, f = _ef(rt, p, frac, Z, idx, cell.T, None, "minimage")
        assert abs(e[0] - ref["energy"]) <= E_RTOL * abs(ref["energy"]), (e[0], ref["energy"])
        assert_force_parity(f[0], ref["forces"], ref64["forces"], where="water 1536")
        print("1536 atoms: kernel-vs-oracle32 %.3f  kernel-vs-fp64 %.3f  oracle32-vs-fp64 %.3f (x tolerance)" % (
            max_force_violation(f[0], ref["forces"]), max_force_violation(f[0], ref64["forces"]),
            max_force_violation(ref["forces"], ref64["forces"])))


def test_unsorted_coo_and_isolated_atoms(rt, oracle):
    go, no = oracle
    pos, Z, cell = syn.water_box(16, 3)
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
    perm = np.random.default_rng(0).permutation(idx.shape[1])
    p = syn.gmnn_params(seed=2)
    frac = pos @ np.linalg.inv(cell)
    e0, f0 = _ef(rt, p, frac, Z, idx, cell.T, off, "offsets")
    e1, f1 = _ef(rt, p, frac, Z, idx[:, perm], cell.T, off[perm], "offsets")
    assert abs(e0[0] - e1[0]) <= 1e-6 * abs(e0[0])
    assert force_tolerance_ok(f1[0], f0[0])
    # no pairs at all: every atom isolated -> forces exactly zero, energy = sum of readout(0)
    empty = np.zeros((2, 0), dtype=np.int32)
    ref = go.energy_forces(p, frac, Z, empty, cell.T, np.zeros((0, 3)), "offsets")
    e2, f2 = _ef(rt, p, frac, Z, empty, cell.T, None, "offsets")
    assert abs(e2[0] - ref["energy"]) <= 1e-6 * max(1.0, abs(ref["energy"]))
    assert np.all(f2 == 0.0)


def test_bitwise_reproducible(rt, oracle):
    go, no = oracle
    pos, Z, cell = syn.water_box(64, 1)
    frac = pos @ np.linalg.inv(cell)
    idx = no.jaxmd_sparse_pairs(frac, cell.T, 6.0)
    p = syn.gmnn_params(seed=9)
    a = _ef(rt, p, frac, Z, idx, cell.T, None, "minimage")
    b = _ef(rt, p, frac, Z, idx, cell.T, None, "minimage")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_properties_at_12288_atoms(rt, oracle):
    """Size-independent properties at BASELINE configs[2]'s size: list symmetry and order, sum F = 0, invariance under a
    rigid translation through the periodic boundary and under atom relabelling.  (E and F against the oracle at this size:
    tests/test_gpu_parity_large.py.)  The transformed evaluations are separate fp32 evaluations, so their forces go through
    the same recorded criterion as every other force comparison, with the fp64 oracle as the truth."""
    go, no = oracle
    pos, Z, cell = syn.water_12288()
    box = cell.T
    frac = pos @ np.linalg.inv(cell)
    n = len(Z)
    cap = int(n * 130)
    idx, n_found, ov = rt.nl_build_minimage(frac, box, 6.5, cap)
    torch.cuda.synchronize()
    nf = int(n_found.item())
    assert not int(ov.item()) and 100 * n < nf < cap
    idx_h = idx[:, :nf].cpu().numpy()
    fwd = idx_h[0].astype(np.int64) * n + idx_h[1]
    rev = idx_h[1].astype(np.int64) * n + idx_h[0]
    assert np.array_equal(np.sort(fwd), np.sort(rev))            # full (symmetric) list
    assert np.all(np.diff(idx_h[1]) >= 0)                        # grouped by central atom, ascending
    same = np.diff(idx_h[1]) == 0
    assert np.all(np.diff(idx_h[0])[same] > 0)                   # neighbours ascending within a row
    assert np.all(idx[:, nf:].cpu().numpy() == n)                # padded with N
    params = syn.gmnn_params(seed=1234)
    dp = rt.DeviceParams(params)
    e, f = rt.energy_forces(dp, frac, Z, idx, box, None, "minimage")
    e, f = e.cpu().numpy(), f.cpu().numpy()[0]
    assert np.isfinite(e).all() and np.isfinite(f).all()
    assert np.abs(f.sum(0)).max() < 1e-3 * np.abs(f).max()
    ref64 = go.energy_forces(params, frac, Z, idx_h, box, None, "minimage", go.fp64_config())
    # rigid translation in fractional space (wraps through the boundary)
    e2, f2 = rt.energy_forces(dp, frac + np.array([0.37, -0.21, 0.05]), Z, idx, box, None, "minimage")
    assert abs(e2.cpu().numpy()[0] - e[0]) <= 1e-6 * abs(e[0])
    assert_force_parity(f2.cpu().numpy()[0], f, ref64["forces"], where="12288 translated vs untranslated")
    # relabel atoms
    perm = np.random.default_rng(1).permutation(n)
    inv = np.argsort(perm)
    idx_p = torch.as_tensor(np.where(idx.cpu().numpy() < n, inv[np.minimum(idx.cpu().numpy(), n - 1)], n).astype(np.int32))
    e3, f3 = rt.energy_forces(dp, frac[perm], Z[perm], idx_p, box, None, "minimage")
    assert abs(e3.cpu().numpy()[0] - e[0]) <= 1e-6 * abs(e[0])
    assert_force_parity(f3.cpu().numpy()[0], f[perm], ref64["forces"][perm], where="12288 relabelled vs original")


@pytest.mark.parametrize("n_contr,use_ntk", [(5, False), (8, True), (3, True)])
def test_truncated_contractions_and_ntk_linear(n_contr, use_ntk):
    """SURVEY.md s8f rank 3 variants: `n_contr < 8` truncates the feature list (gaussian_moment_descriptor.py:101) and
    `use_ntk` rescales the dense layers (ntk_linear.py:42-45); both against the oracle on a multi-image water cell."""
    from apax_b200 import runtime as rt
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no

    pos, Z, cell = syn.water_box(16, 4)
    idx, offsets = no.vesin_idx_offsets(pos, cell, 6.0)
    frac = pos @ np.linalg.inv(cell)
    n_feat = rt.FEATURE_COUNTS[n_contr - 1]
    params = syn.gmnn_params(seed=5, n_features=n_feat, random_bias=True, random_scale_shift=True)
    cfg = go.GMNNConfig(n_contr=n_contr, use_ntk=use_ntk)
    ref = go.energy_forces(params, frac, Z, idx, cell.T, offsets, "offsets", cfg)
    ref64 = go.energy_forces(params, frac, Z, idx, cell.T, offsets, "offsets", go.fp64_config(cfg))
    dp = rt.DeviceParams(params, {"n_contr": n_contr, "use_ntk": use_ntk})
    e, f = rt.energy_forces(dp, frac, Z, idx, cell.T, offsets, "offsets")
    torch.cuda.synchronize()
    e, f = e.cpu().numpy(), f.cpu().numpy()[0]
    assert abs(e[0] - ref["energy"]) <= 1e-6 * abs(ref["energy"])
    assert_force_parity(f, ref["forces"], ref64["forces"], where=f"n_contr={n_contr} ntk={use_ntk}")
    g = rt.descriptor(dp, frac, Z, idx, cell.T, offsets, "offsets").cpu().numpy()
    g_ref = go.descriptor_only(params, frac, Z, idx, cell.T, offsets, "offsets", cfg)["g"]
    assert g.shape == (len(Z), n_feat)
    assert np.abs(g - g_ref).max() <= 2e-6 * np.abs(g_ref).max()


def test_stress_times_vol_min_image(golden):
    """stress x volume on the min-image path against the fixture written by the reference's own
    stress_times_vol / EnergyDerivativeModel(calc_stress=True) source, and against the oracle in fp64."""
    from apax_b200.nn.builder import GMNNBuilder
    from oracle import gmnn_oracle as go

    d = golden("water_stress")
    params = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True)
    model = GMNNBuilder({"calc_stress": True}).build_energy_derivative_model(init_box=d["box"], inference_disp_fn=object())
    out = model.apply(params, d["frac"], d["Z"], d["idx"], d["box"], None)
    torch.cuda.synchronize()
    s = out["stress"].cpu().numpy()
    ref64 = go.stress_times_vol(params, d["frac"], d["Z"], d["idx"], d["box"], None, "minimage", go.fp64_config())
    scale = np.abs(ref64).max()
    assert np.abs(s - d["stress"]).max() <= 2e-6 * scale                      # vs the reference source (fp32 model)
    assert np.abs(s - ref64).max() <= 2.0 * np.abs(d["stress"] - ref64).max() + 1e-6 * scale
    assert abs(float(out["energy"].item()) - float(d["energy"])) <= 1e-6 * abs(float(d["energy"]))


def test_empirical_corrections_zbl_and_exponential(golden):
    """ZBLRepulsion + ExponentialRepulsion in EnergyModel.corrections against the fixture written by
    the reference's own apax/layers/empirical.py, and against the oracle in fp64."""
    from apax_b200.nn.builder import GMNNBuilder
    from oracle import gmnn_oracle as go

    d = golden("water_corrections")
    params = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True, random_scale_shift=True)
    em = params["params"]["energy_model"]
    em["corrections_0"] = {"coefficients": d["zbl_coefficients"], "rep_scale": d["zbl_rep_scale"]}
    em["corrections_1"] = {"rep_scale": d["exp_rep_scale"], "rep_prefactor": d["exp_rep_prefactor"]}
    cfg = {"empirical_corrections": [{"name": "zbl", "r_max": float(d["zbl_r_max"])}, {"name": "exponential", "r_max": float(d["exp_r_max"])}]}
    model = GMNNBuilder(cfg).build_energy_derivative_model(init_box=d["cell"].T)
    out = model.apply(params, d["frac"], d["Z"], d["idx"], d["cell"], d["offsets"])
    torch.cuda.synchronize()
    e, f = float(out["energy"].item()), out["forces"].cpu().numpy()
    ocfg = go.GMNNConfig(zbl_r_max=float(d["zbl_r_max"]), exp_r_max=float(d["exp_r_max"]))
    ref64 = go.energy_forces(params, d["frac"], d["Z"], d["idx"], d["cell"], d["offsets"], "offsets", go.fp64_config(ocfg))
    assert abs(e - float(d["energy"])) <= 1e-6 * abs(float(d["energy"]))
    assert_force_parity(f, d["forces"], ref64["forces"], where="corrections", ulp_floor=2.0)   # |F| up to 59 eV/A: 1 ulp = 3.8e-6
    plain = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True, random_scale_shift=True)
    e0 = GMNNBuilder().build_energy_derivative_model(init_box=d["cell"].T).apply(plain, d["frac"], d["Z"], d["idx"], d["cell"], d["offsets"])
    assert abs(float(e0["energy"].item()) - e) > 1.0      # the corrections are a large part of this energy


def test_stress_times_vol_offsets_branch(golden):
    """calc_stress on the vesin / offsets path: dU/d(eps) at doubled displacements, as the reference's disp_fn computes it
    (apax/layers/distances.py:12-22), against the fixture written by the reference's own source and the oracle in fp64."""
    from apax_b200.nn.builder import GMNNBuilder
    from oracle import gmnn_oracle as go

    d = golden("water_stress_offsets")
    params = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True)
    model = GMNNBuilder({"calc_stress": True}).build_energy_derivative_model(init_box=d["cell"].T)
    out = model.apply(params, d["frac"], d["Z"], d["idx"], d["cell"], d["offsets"])
    torch.cuda.synchronize()
    s = out["stress"].cpu().numpy()
    ref64 = go.stress_times_vol(params, d["frac"], d["Z"], d["idx"], d["cell"], d["offsets"], "offsets", go.fp64_config())
    scale = np.abs(d["stress"]).max()
    assert np.abs(s - d["stress"]).max() <= 2e-6 * scale                      # vs the reference source (fp32 model)
    assert np.abs(s - ref64).max() <= 2.0 * np.abs(d["stress"] - ref64).max() + 1e-6 * scale
    # energy and forces of the same call are those of the unperturbed evaluation
    assert abs(float(out["energy"].item()) - float(d["energy"])) <= 1e-6 * abs(float(d["energy"]))
    assert_force_parity(out["forces"].cpu().numpy(), d["forces"], where="stress_offsets")


def _with_corrections(params, seed=3):
    """ZBL + exponential blocks as flax names them (corrections_0, corrections_1); ranges that catch water's O-H and H-H pairs."""
    rng = np.random.default_rng(seed)
    em = params["params"]["energy_model"]
    em["corrections_0"] = {"coefficients": rng.normal(size=(4, 1)).astype(np.float32), "rep_scale": np.array([-1.2], dtype=np.float32)}
    em["corrections_1"] = {"rep_scale": (rng.uniform(0.3, 1.2, size=119) * rng.choice([-1, 1], size=119)).astype(np.float32),
                           "rep_prefactor": rng.uniform(1.0, 20.0, size=119).astype(np.float32)}
    return params


def test_corrections_min_image_energy_forces_stress_vs_oracle(rt, oracle):
    """The pair corrections on the JaxMD min-image path (E, F and stress x volume share the corrected per-pair records) against
    the oracle in reference precision and in fp64; 768 atoms, list with skin."""
    go, no = oracle
    pos, Z, cell = syn.water_box(256, 11)
    frac = pos @ np.linalg.inv(cell)
    idx, occ, ov = no.jaxmd_sparse_list(frac, cell.T, 6.0, 0.5)
    assert not ov
    p = _with_corrections(syn.gmnn_params(seed=99, random_scale_shift=True, random_bias=True))
    kcfg, ocfg = {"zbl_r_max": 1.8, "exp_r_max": 1.6}, go.GMNNConfig(zbl_r_max=1.8, exp_r_max=1.6)
    ref = go.energy_forces(p, frac, Z, idx, cell.T, None, "minimage", ocfg)
    ref64 = go.energy_forces(p, frac, Z, idx, cell.T, None, "minimage", go.fp64_config(ocfg))
    dp = rt.DeviceParams(p, kcfg)
    e, f, s = rt.energy_forces(dp, frac, Z, idx, cell.T, None, "minimage", stress=True)
    torch.cuda.synchronize()
    e, f, s = e.cpu().numpy(), f.cpu().numpy(), s.cpu().numpy()
    assert abs(e[0] - ref["energy"]) <= E_RTOL * abs(ref["energy"]), (e[0], ref["energy"])
    # the per-pair records are fp32 (scalar dE/dr / r and the displacement): two roundings of a 230 eV/A O-H pair force show up
    # as a few 1e-5 eV/A in a net force of 75 eV/A -- 20x inside the literal tolerance (checked above all else), a few ulps of |F|max
    assert_force_parity(f[0], ref["forces"], ref64["forces"], where="corrections min-image", ulp_floor=8.0)
    st32 = go.stress_times_vol(p, frac, Z, idx, cell.T, None, "minimage", ocfg)
    st64 = go.stress_times_vol(p, frac, Z, idx, cell.T, None, "minimage", go.fp64_config(ocfg))
    scale = np.abs(st64).max()
    assert np.abs(s - st32).max() <= 2e-6 * scale
    assert np.abs(s - st64).max() <= 2.0 * np.abs(st32 - st64).max() + 1e-6 * scale


def test_corrections_with_shallow_ensemble_heads(rt, oracle):
    """Every head of a shallow ensemble carries the same correction energy and forces (EnergyModel adds it after the readout)."""
    go, no = oracle
    pos, Z, cell = syn.water_box(32, 5)
    idx, offsets = no.vesin_idx_offsets(pos, cell, 6.0)
    frac = pos @ np.linalg.inv(cell)
    n_ens = 4
    p = _with_corrections(syn.gmnn_params(seed=5, n_out=n_ens, random_scale_shift=True, random_bias=True))
    kcfg, ocfg = {"zbl_r_max": 1.8, "exp_r_max": 1.6}, go.GMNNConfig(zbl_r_max=1.8, exp_r_max=1.6)
    e, f = _ef(rt, p, frac, Z, idx, cell, offsets, "offsets", kcfg)
    r32 = go.shallow_ensemble(p, frac, Z, idx, cell, offsets, "offsets", ocfg)
    r64 = go.shallow_ensemble(p, frac, Z, idx, cell, offsets, "offsets", go.fp64_config(ocfg))
    assert_energy_parity(e, r32["energy_ensemble"], r64["energy_ensemble"], where="corrections ensemble")
    assert_force_parity(np.transpose(f, (1, 2, 0)), r32["forces_ensemble"], r64["forces_ensemble"], where="corrections ensemble",
                        ulp_floor=8.0)
