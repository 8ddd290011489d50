"""The reference-facing host mirrors (same names / call shapes as apax) driven end to end on the GPU."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn
from conftest import assert_force_parity

pytestmark = pytest.mark.gpu


def test_descriptor_module_apply(golden):
    from apax_b200.layers.descriptor import GaussianMomentDescriptor

    d = golden("descriptor_unit")
    p = syn.gmnn_params(seed=int(d["param_seed"]))
    emb = p["params"]["energy_model"]["representation"]["radial_fn"]["atomic_type_embedding"]
    g = GaussianMomentDescriptor().apply({"params": {"radial_fn": {"atomic_type_embedding": emb}}}, d["dR"], d["Z"], d["idx"])
    assert tuple(g.shape) == (3, 360)  # tests/unit_tests/layers/descriptor/test_gaussian_moment_descriptor.py:30
    assert np.abs(g.cpu().numpy() - d["g"]).max() <= 2e-6 * np.abs(d["g"]).max()


def test_builder_models_and_neighbor_provider(golden):
    """The JaxMD path of ASECalculator (ase_calc.py:39-73, 530-542) with the mirrored names."""
    from apax_b200.nn.builder import GMNNBuilder
    from apax_b200.utils.jax_md_reduced import partition

    d = golden("water_min_image")
    params = syn.gmnn_params(seed=int(d["param_seed"]))
    box = d["box"]
    disp_fn = object()  # only its presence selects the min-image branch (distances.py:41-46)
    neighbor_fn = partition.neighbor_list(disp_fn, box, 6.0, float(d["skin"]), fractional_coordinates=True,
                                          disable_cell_list=True, format=partition.Sparse)
    nbrs = neighbor_fn.allocate(d["frac"], box=box)
    assert nbrs.idx.shape == d["idx"].shape and np.array_equal(nbrs.idx.cpu().numpy(), d["idx"])
    assert not nbrs.did_buffer_overflow
    same = nbrs.update(d["moved_small"], box=box)
    assert same is nbrs                                               # inside the skin: list kept
    moved = nbrs.update(d["moved_big"], box=box)
    assert np.array_equal(moved.idx.cpu().numpy(), d["idx_big"])       # rebuilt, same capacity
    model = GMNNBuilder().build_energy_derivative_model(init_box=box, inference_disp_fn=disp_fn)
    out = model.apply(params, d["frac"], d["Z"], nbrs, box, np.zeros((nbrs.idx.shape[1], 3)))
    assert abs(float(out["energy"]) - d["energy"]) <= 1e-6 * abs(d["energy"])
    f = out["forces"].cpu().numpy()
    from oracle import gmnn_oracle as go
    ref64 = go.energy_forces(params, d["frac"], d["Z"], d["idx"], box, None, "minimage", go.fp64_config())
    assert_force_parity(f, d["forces"], ref64["forces"], where="builder model, golden min-image")
    e_only, props = model.energy_model.apply(params, d["frac"], d["Z"], nbrs, box, None)
    assert props == {} and abs(float(e_only) - float(out["energy"])) <= 1e-9 * abs(d["energy"])


def test_shallow_ensemble_model(golden):
    from apax_b200.nn.builder import GMNNBuilder

    d = golden("water_ensemble")
    n_ens = int(d["n_ens"])
    params = syn.gmnn_params(seed=int(d["param_seed"]), n_out=n_ens, random_scale_shift=True, random_bias=True)
    cfg = {"ensemble": {"kind": "shallow", "n_members": n_ens, "force_variance": True, "chunk_size": None}}
    model = GMNNBuilder(cfg).build_energy_derivative_model(init_box=d["cell"].T)
    out = model.apply(params, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"])
    from conftest import assert_energy_parity
    from oracle import gmnn_oracle as go

    r64 = go.shallow_ensemble(params, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"], "offsets", go.fp64_config())
    assert_energy_parity(out["energy_ensemble"].cpu().numpy(), d["energy_ensemble"], r64["energy_ensemble"], where="ShallowEnsembleModel")
    assert_energy_parity(out["energy"].cpu().numpy(), d["energy"], r64["energy"], where="ShallowEnsembleModel mean")
    # the spread of the heads is a difference of energies: its error scales with the energies, not with itself
    e_scale = np.abs(d["energy_ensemble"]).max()
    assert abs(float(out["energy_uncertainty"]) - float(d["energy_uncertainty"])) <= 1e-6 * e_scale
    for k in ("forces", "forces_ensemble"):
        assert_force_parity(out[k].cpu().numpy(), d[k], r64[k], where=f"ShallowEnsembleModel {k}")
    f_scale = np.abs(d["forces_ensemble"]).max(axis=-1)
    assert np.all(np.abs(out["forces_uncertainty"].cpu().numpy() - d["forces_uncertainty"]) <= 1e-5 * f_scale + 1e-6)


def test_ase_calculator_mirror():
    from apax_b200.md.ase_calc import ASECalculator, Atoms
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no

    params = syn.gmnn_params(seed=1234)
    calc = ASECalculator(params, dr_threshold=0.5)
    pos, Z, cell = syn.water_96()                       # BASELINE config 1: vesin (multi-image) path
    res = calc.calculate(Atoms(pos, Z, cell))
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
    ref = go.energy_forces(params, pos @ np.linalg.inv(cell), Z, idx, cell.T, off, "offsets")
    assert abs(res["energy"] - ref["energy"]) <= 1e-6 * abs(ref["energy"])
    assert_force_parity(res["forces"], ref["forces"], where="ASECalculator 96")
    assert res["forces"].dtype == np.float64 and isinstance(res["energy"], float)
    # capacity overflow -> transparent re-initialisation (ase_calc.py:381-384)
    small = ASECalculator(params, max_pairs_per_atom=20)
    res2 = small.calculate(Atoms(pos, Z, cell))
    assert abs(res2["energy"] - res["energy"]) <= 1e-12 * abs(res["energy"])


def test_energy_fn_for_md():
    from apax_b200.md.simulate import create_energy_fn
    from apax_b200.nn.builder import GMNNBuilder
    from apax_b200.utils.jax_md_reduced import partition

    pos, Z, cell = syn.water_box(64, 3)
    box = cell.T.copy()
    frac = pos @ np.linalg.inv(cell)
    params = syn.gmnn_params(seed=4)
    model = GMNNBuilder().build_energy_model(init_box=box, inference_disp_fn=object())
    nbrs = partition.neighbor_list(None, box, 6.0, 0.5, fractional_coordinates=True, format=partition.Sparse).allocate(frac, box=box)
    energy_fn = create_energy_fn(model, params, Z, n_models=1)
    e = float(energy_fn(frac, nbrs, box))
    f = energy_fn.force_fn(frac, nbrs, box).cpu().numpy()
    h = 5e-3  # fp32 energy noise ~2e-5 eV: the difference quotient is only good to ~1e-2
    inv = np.linalg.inv(cell)
    for a, k in [(0, 0), (17, 2)]:
        dc = np.zeros(3)
        dc[k] = h
        fp, fm = frac.copy(), frac.copy()
        fp[a] += dc @ inv
        fm[a] -= dc @ inv
        fd = -(float(energy_fn(fp, nbrs, box)) - float(energy_fn(fm, nbrs, box))) / (2 * h)
        assert abs(fd - f[a, k]) < 2e-2 * max(1.0, abs(f[a, k]))


def test_calculator_stress_both_neighbour_paths():
    """`stress` through the calculator (apax_ctx_stress): EnergyDerivativeModel(calc_stress) followed by ProcessStress
    (apax/md/function_transformations.py:140-159: val.T / det(box)), on the min-image path and on the image-list path (where
    the reference differentiates at doubled displacements), against the oracle's stress_times_vol."""
    from apax_b200.md.ase_calc import ASECalculator, Atoms
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no

    params = syn.gmnn_params(seed=8, random_bias=True)
    calc = ASECalculator(params, model_config={"calc_stress": True}, dr_threshold=0.5)
    assert "stress" in calc.implemented_properties
    for n_mol, seed, mode in ((120, 2, "minimage"), (32, 0, "offsets")):
        pos, Z, cell = syn.water_box(n_mol, seed)
        res = calc.calculate(Atoms(pos, Z, cell))
        frac = pos @ np.linalg.inv(cell)
        if mode == "minimage":
            idx, off, box = no.jaxmd_sparse_pairs(frac, cell.T, 6.0), None, cell.T
        else:
            idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
            box = cell                                                   # the vesin path hands box = atoms.cell.array (ase_calc.py:388-390)
            frac = pos @ np.linalg.inv(cell)
        st32 = go.stress_times_vol(params, frac if mode == "minimage" else (np.linalg.inv(box) @ pos.T).T, Z, idx, box, off, mode)
        st64 = go.stress_times_vol(params, frac if mode == "minimage" else (np.linalg.inv(box) @ pos.T).T, Z, idx, box, off, mode, go.fp64_config())
        vol = np.linalg.det(cell)
        ref32, ref64 = st32.T / vol, st64.T / vol
        scale = np.abs(ref64).max()
        assert res["stress"].shape == (3, 3)
        assert np.abs(res["stress"] - ref32).max() <= 4e-6 * scale, (mode, np.abs(res["stress"] - ref32).max() / scale)
        assert np.abs(res["stress"] - ref64).max() <= 2.0 * np.abs(ref32 - ref64).max() + 1e-6 * scale, mode
    plain = ASECalculator(params)
    assert "stress" not in plain.calculate(Atoms(pos, Z, cell))


def test_changed_atomic_numbers_reprepare_the_list():
    """A change of `numbers` at fixed N, cell and within-skin positions must not reuse the compact species table of the old
    list (the reference re-initialises on any change of numbers, apax/md/ase_calc.py:366-370)."""
    from apax_b200 import runtime as rt
    from apax_b200.md.ase_calc import ASECalculator, Atoms

    pos, Z, cell = syn.water_box(120, 2)                                 # min-image path: the list survives between calls
    params = syn.gmnn_params(seed=8)
    Z2 = Z.copy()
    h = np.nonzero(Z == 1)[0]
    Z2[h[::7]] = 9                                                        # a species that was absent from the first list
    dp = rt.DeviceParams(params)
    ctx = rt.HostContext(dp, len(Z), 130 * len(Z))
    ctx.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.5)
    e_reused, f_reused = ctx.calculate(pos, Z2.astype(np.int32), cell, dr_threshold=0.5)      # same positions: skin test says "keep"
    fresh = rt.HostContext(dp, len(Z), 130 * len(Z))
    e_fresh, f_fresh = fresh.calculate(pos, Z2.astype(np.int32), cell, dr_threshold=0.5)
    assert np.array_equal(e_reused, e_fresh) and np.array_equal(f_reused, f_fresh)
    e_old, _ = fresh.calculate(pos, Z.astype(np.int32), cell, dr_threshold=0.5)
    assert abs(e_old[0] - e_fresh[0]) > 1e-3                              # the species really matter
    calc = ASECalculator(params, dr_threshold=0.5)
    r1 = calc.calculate(Atoms(pos, Z, cell))
    r2 = calc.calculate(Atoms(pos, Z2, cell))
    assert abs(r2["energy"] - e_fresh[0]) <= 1e-12 * abs(e_fresh[0]) and r1["energy"] != r2["energy"]


def test_mean_forces_single_backward_pass_matches_per_head_mean(golden):
    """ShallowEnsembleModel(force_variance=False) (apax/nn/models.py:265-267) and the MD energy function of a shallow ensemble
    (apax/md/simulate.py:55-58) need only -grad(mean energy): one backward pass with the mean head's output weights
    (APAX_OPT_MEAN_FORCES) against the mean of the per-head forces and against the reference-source fixture."""
    from apax_b200 import runtime as rt
    from apax_b200.nn.builder import GMNNBuilder
    from oracle import gmnn_oracle as go

    d = golden("water_ensemble")
    n_ens = int(d["n_ens"])
    params = syn.gmnn_params(seed=int(d["param_seed"]), n_out=n_ens, random_scale_shift=True, random_bias=True)
    cfg = {"ensemble": {"kind": "shallow", "n_members": n_ens, "force_variance": False}}
    model = GMNNBuilder(cfg).build_energy_derivative_model(init_box=d["cell"].T)
    c0 = rt.lib().apax_kernel_launch_count()
    out = model.apply(params, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"])
    torch.cuda.synchronize()
    launches_mean = rt.lib().apax_kernel_launch_count() - c0
    assert set(out) == {"energy", "energy_ensemble", "energy_uncertainty", "forces"}
    r64 = go.shallow_ensemble(params, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"], "offsets", go.fp64_config())
    assert_force_parity(out["forces"].cpu().numpy(), d["forces"], r64["forces"], where="ShallowEnsembleModel mean forces (one pass)")
    assert np.allclose(out["energy_ensemble"].cpu().numpy(), d["energy_ensemble"], rtol=1e-6, atol=0)
    full = GMNNBuilder({"ensemble": {"kind": "shallow", "n_members": n_ens, "force_variance": True}}).build_energy_derivative_model(init_box=d["cell"].T)
    c0 = rt.lib().apax_kernel_launch_count()
    ref = full.apply(params, d["frac"], d["Z"], d["idx"], d["box"], d["offsets"])
    torch.cuda.synchronize()
    launches_full = rt.lib().apax_kernel_launch_count() - c0
    assert launches_full > launches_mean + 3 * (n_ens - 1)           # the per-head passes really are gone
    fm, ff = out["forces"].cpu().numpy(), ref["forces"].cpu().numpy()
    assert np.all(np.abs(fm - ff) <= 1e-5 * np.abs(ff) + 1e-6)
