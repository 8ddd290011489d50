"""Generate golden fixtures by executing the reference's OWN hot-path source files
(/root/reference/apax/...) under the torch-backed jax/flax shim in jax_shim.py.

Run by hand in the build container:  python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The GPU box has no /root/reference; tests only read the .npz.

Modules executed unmodified from the reference: apax/layers/{distances,masking,readout,
ntk_linear,activation,scaling}.py, apax/layers/descriptor/{gaussian_moment_descriptor,
basis_functions,moments,triangular_indices}.py, apax/nn/models.py, apax/utils/math.py,
apax/utils/jax_md_reduced/{space,partition,util,dataclasses}.py.
vesin is absent, so multi-image lists come from oracle/neighbor_oracle.py (an INPUT to the
reference model here, not something being checked).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import jax_shim  # noqa: E402

jax_shim.install("/root/reference")

from apax.layers.activation import variance_preserving_swish  # noqa: E402
from apax.layers.descriptor.basis_functions import GaussianBasis, RadialFunction  # noqa: E402
from apax.layers.descriptor.gaussian_moment_descriptor import GaussianMomentDescriptor  # noqa: E402
from apax.layers.readout import AtomisticReadout  # noqa: E402
from apax.layers.scaling import PerElementScaleShift  # noqa: E402
from apax.nn.models import EnergyDerivativeModel, EnergyModel, ShallowEnsembleModel  # noqa: E402
from apax.utils.jax_md_reduced import partition, space  # noqa: E402

from apax_b200 import synthetic as syn  # noqa: E402
from oracle import neighbor_oracle as no  # noqa: E402


def to_t(tree):
    if isinstance(tree, dict):
        return {k: to_t(v) for k, v in tree.items()}
    return torch.as_tensor(np.asarray(tree))


def build_model(init_box, inference_disp_fn=None, n_ens=0, scale=1.0, shift=0.0, apply_mask=True,
                mask_atoms=True, ensemble=False):
    """What GMNNBuilder.build_energy_derivative_model assembles (apax/nn/builder.py:160-260) with the
    model_config.py defaults named in SURVEY.md §8."""
    descriptor = GaussianMomentDescriptor(
        radial_fn=RadialFunction(n_radial=5, basis_fn=GaussianBasis(n_basis=7, r_min=0.5, r_max=6.0, dtype="fp32"),
                                 n_species=119, emb_init="uniform", use_embed_norm=True, dtype="fp32"),
        n_contr=8, dtype="fp32", apply_mask=apply_mask)
    readout = AtomisticReadout(units=[256, 256], activation_fn=variance_preserving_swish, w_init="lecun",
                               b_init="zeros", use_ntk=False, n_shallow_ensemble=n_ens, dtype="fp32")
    scale_shift = PerElementScaleShift(n_species=119, scale=scale, shift=shift, dtype="fp64")
    em = EnergyModel(representation=descriptor, readout=readout, scale_shift=scale_shift, init_box=init_box,
                     inference_disp_fn=inference_disp_fn, mask_atoms=mask_atoms)
    if ensemble:
        return ShallowEnsembleModel(em, force_variance=True)
    return EnergyDerivativeModel(em)


def np_out(res):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in res.items()}


def case_descriptor_unit():
    """Geometry of tests/unit_tests/layers/descriptor/test_gaussian_moment_descriptor.py:8-31."""
    dR = np.array([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [-1, 1, 0], [0, -1, 0], [1, -1, 0]], dtype=np.float32)
    Z = np.array([8, 1, 1])
    idx = np.array([[1, 2, 0, 2, 0, 1], [0, 0, 1, 1, 2, 2]])
    params = syn.gmnn_params(seed=7)
    desc = GaussianMomentDescriptor()
    emb = params["params"]["energy_model"]["representation"]["radial_fn"]["atomic_type_embedding"]
    g = desc.apply({"params": {"radial_fn": {"atomic_type_embedding": torch.as_tensor(emb)}}},
                   torch.as_tensor(dR), torch.as_tensor(Z), torch.as_tensor(idx))
    assert tuple(g.shape) == (3, 360)
    np.savez(os.path.join(HERE, "descriptor_unit.npz"), dR=dR, Z=Z, idx=idx, param_seed=7, g=g.numpy())


def case_gas_padded():
    """Geometry + scale/shift of tests/unit_tests/model/test_apax.py:10-62 (padding invariance)."""
    R = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    Z = np.array([1, 2, 2])
    idx = np.array([[1, 2, 0, 2, 0, 1], [0, 0, 1, 1, 2, 2]])
    box = np.zeros(3)
    shift = np.zeros(119)
    shift[:3] = [0.0, 100.0, 200.0]
    scale = np.ones(119)
    scale[:3] = [1.0, 1.2, 1.6]
    params = syn.gmnn_params(seed=11, scale=scale, shift=shift)
    tparams = to_t(params)
    offsets = np.zeros((6, 3))
    m = build_model(box, apply_mask=False, mask_atoms=False, scale=scale[:, None], shift=shift)
    res = np_out(m.apply(tparams, torch.as_tensor(R), torch.as_tensor(Z), torch.as_tensor(idx), torch.as_tensor(box),
                         torch.as_tensor(offsets)))
    Rp = np.concatenate([R, np.zeros((1, 3))])
    Zp = np.concatenate([Z, [0]])
    idxp = np.concatenate([idx, [[0], [0]]], axis=1)
    offp = np.zeros((7, 3))
    mp = build_model(box, apply_mask=True, mask_atoms=True, scale=scale[:, None], shift=shift)
    resp = np_out(mp.apply(tparams, torch.as_tensor(Rp), torch.as_tensor(Zp), torch.as_tensor(idxp),
                           torch.as_tensor(box), torch.as_tensor(offp)))
    assert abs(res["energy"] - resp["energy"]) < 1e-6
    assert np.all(np.abs(res["forces"] - resp["forces"][:-1]) < 1e-6)
    np.savez(os.path.join(HERE, "gas_padded.npz"), R=R, Z=Z, idx=idx, Rp=Rp, Zp=Zp, idxp=idxp, scale=scale, shift=shift,
             param_seed=11, energy=res["energy"], forces=res["forces"], energy_padded=resp["energy"],
             forces_padded=resp["forces"])


def case_multi_image(n_mol=12, seed=3):
    """Small periodic water box (L < 2 r_max): the vesin path of ASECalculator.calculate
    (apax/md/ase_calc.py:386-390): frac positions, box = cell, disp_fn + Cartesian offsets."""
    pos, Z, cell = syn.water_box(n_mol, seed)
    idx, offsets = no.vesin_idx_offsets(pos, cell, 6.0)
    pad = int(idx.shape[1] * 1.5)
    idx, offsets = no.vesin_idx_offsets(pos, cell, 6.0, padded_length=pad)
    box = cell  # ase_calc.py:359,389: box = atoms.cell.array (not transposed on this path)
    frac = np.asarray(space.transform(torch.as_tensor(np.linalg.inv(box)), torch.as_tensor(pos)))
    params = syn.gmnn_params(seed=1234)
    m = build_model(np.array(cell.T))
    res = np_out(m.apply(to_t(params), torch.as_tensor(frac), torch.as_tensor(Z.astype(np.int64)),
                         torch.as_tensor(idx.astype(np.int64)), torch.as_tensor(box), torch.as_tensor(offsets)))
    np.savez(os.path.join(HERE, "water_multi_image.npz"), positions=pos, frac=frac, Z=Z, cell=cell, idx=idx,
             offsets=offsets, box=box, param_seed=1234, energy=res["energy"], forces=res["forces"])


def case_min_image(n_mol=64, seed=5, skin=0.5):
    """Periodic water box with L > 2 r_max: the JaxMD path (ase_calc.py:39-73, 530-542):
    partition.neighbor_list(Sparse, fractional) + periodic_general displacement, zero offsets."""
    pos, Z, cell = syn.water_box(n_mol, seed)
    box = torch.as_tensor(cell.T.copy())
    disp_fn, _ = space.periodic_general(box, fractional_coordinates=True)
    nfn = partition.neighbor_list(disp_fn, box, 6.0, skin, fractional_coordinates=True, disable_cell_list=True,
                                  format=partition.Sparse)
    frac = space.transform(torch.linalg.inv(box), torch.as_tensor(pos))
    nbrs = nfn.allocate(frac, box=box)
    idx = nbrs.idx.numpy()
    n_valid = int((idx[0] < len(Z)).sum())
    # skin test on slightly moved positions (partition.py:771-779)
    rng = np.random.default_rng(0)
    moved_small = frac + torch.as_tensor(rng.normal(scale=0.002, size=frac.shape)) @ torch.linalg.inv(box).T
    moved_big = frac.clone()
    moved_big[7] += torch.as_tensor([0.0, 0.3, 0.0], dtype=torch.float64) @ torch.linalg.inv(box).T
    same_small = nbrs.update(moved_small, box=box)
    same_big = nbrs.update(moved_big, box=box)
    rebuilt_small = not torch.equal(same_small.reference_position, nbrs.reference_position)
    rebuilt_big = not torch.equal(same_big.reference_position, nbrs.reference_position)
    params = syn.gmnn_params(seed=1234)
    m = build_model(np.array(cell.T), inference_disp_fn=disp_fn)
    offsets = torch.zeros((idx.shape[1], 3), dtype=torch.int64)  # jnp.full([P,3], 0) (ase_calc.py:540)
    res = np_out(m.apply(to_t(params), frac, torch.as_tensor(Z.astype(np.int64)), nbrs.idx.long(), box, offsets))
    np.savez(os.path.join(HERE, "water_min_image.npz"), positions=pos, frac=frac.numpy(), Z=Z, cell=cell,
             box=box.numpy(), idx=idx, n_valid=n_valid, skin=skin, moved_small=moved_small.numpy(),
             moved_big=moved_big.numpy(), rebuilt_small=rebuilt_small, rebuilt_big=rebuilt_big,
             overflow=int(nbrs.did_buffer_overflow), param_seed=1234, energy=res["energy"], forces=res["forces"],
             idx_big=same_big.idx.numpy())


def case_ensemble(n_mol=10, seed=9, n_ens=4):
    """Shallow ensemble (apax/nn/models.py:199-275) on a small multi-image water cell."""
    pos, Z, cell = syn.water_box(n_mol, seed)
    idx, offsets = no.vesin_idx_offsets(pos, cell, 6.0)
    box = cell
    frac = pos @ np.linalg.inv(box).T
    frac = np.asarray(space.transform(torch.as_tensor(np.linalg.inv(box)), torch.as_tensor(pos)))
    params = syn.gmnn_params(seed=77, n_out=n_ens, random_scale_shift=True, random_bias=True)
    em = params["params"]["energy_model"]
    m = build_model(np.array(cell.T), n_ens=n_ens, ensemble=True,
                    scale=em["scale_shift"]["scale_per_element"], shift=em["scale_shift"]["shift_per_element"])
    res = np_out(m.apply(to_t(params), torch.as_tensor(frac), torch.as_tensor(Z.astype(np.int64)),
                         torch.as_tensor(idx.astype(np.int64)), torch.as_tensor(box), torch.as_tensor(offsets)))
    np.savez(os.path.join(HERE, "water_ensemble.npz"), positions=pos, frac=frac, Z=Z, cell=cell, idx=idx, offsets=offsets,
             box=box, param_seed=77, n_ens=n_ens, **res)


def case_md_nvt(n=24, seed=2):
    """jax_md's nvt_nose_hoover.apply_fn (apax/utils/jax_md_reduced/simulate.py:611-637) as apax runs it
    (apax/md/simulate.py:100-119), driven by a harmonic tether force so that only the integrator is exercised."""
    from apax.utils.jax_md_reduced import simulate

    rng = np.random.default_rng(seed)
    cell = np.array([[9.0, 0.0, 0.0], [1.0, 8.0, 0.0], [0.5, -0.7, 10.0]])
    box = torch.as_tensor(cell.T.copy())
    frac0 = rng.uniform(size=(n, 3))
    anchor = torch.as_tensor(frac0.copy())
    mass = rng.uniform(1.0, 16.0, size=n)
    kT, dt = 0.3, 0.02
    p0 = rng.normal(size=(n, 3)) * np.sqrt(mass * kT)[:, None]
    p0 -= p0.mean(0)
    kspring = 0.7

    def force_fn(R, **kwargs):
        d = R - anchor
        d = d - torch.round(d)
        return -kspring * (d @ box.T)            # Cartesian restoring force

    _, shift_fn = space.periodic_general(box, fractional_coordinates=True)
    init_fn, apply_fn = simulate.nvt_nose_hoover(force_fn, shift_fn, dt, kT)   # defaults: chain 5 / 2 / 3, tau = 100 dt
    R = torch.as_tensor(frac0.copy())
    thermostat = simulate.nose_hoover_chain(simulate.f32(dt), 5, 2, 3, simulate.f32(simulate.f32(dt) * 100))
    state = simulate.NVTNoseHooverState(R, torch.as_tensor(p0.copy()), force_fn(R), torch.as_tensor(mass.copy())[:, None], None)
    KE = simulate.kinetic_energy(state)
    state = state.set(chain=thermostat.initialize(3 * n, KE, kT))
    snaps = {}
    for step in range(1, 8):
        state = apply_fn(state)
        if step in (1, 7):
            snaps[f"position_{step}"] = state.position.numpy().copy()
            snaps[f"momentum_{step}"] = state.momentum.numpy().copy()
            snaps[f"xi_{step}"] = state.chain.position.numpy().copy()
            snaps[f"p_xi_{step}"] = state.chain.momentum.numpy().copy()
            snaps[f"Q_{step}"] = state.chain.mass.numpy().astype(np.float64)
            snaps[f"KE_{step}"] = float(state.chain.kinetic_energy)
    np.savez(os.path.join(HERE, "md_nvt.npz"), cell=cell, frac0=frac0, mass=mass, p0=p0, kT=kT, dt=dt, kspring=kspring,
             KE0=float(KE), **snaps)


def case_train_step(n_struct=3, seed=21):
    """Loss and parameter gradients of one training batch, from the reference's own `Loss` / `LossCollection`
    (apax/train/loss.py:199-259) applied to the reference's own EnergyDerivativeModel (apax/nn/models.py:137-171)
    the way calc_loss does (apax/train/trainer.py:253-263).  The jax.vmap over structures (apax/train/run.py:204) is
    written as a loop; jax.value_and_grad w.r.t. the parameters (trainer.py:332) is torch.autograd.grad over a graph
    that contains the inner value_and_grad (shim HIGHER_ORDER)."""
    from apax.train.loss import Loss, LossCollection

    from oracle import train_oracle as tor

    inputs, labels = tor.ethanol_training_batch(n_struct, seed=seed, pad_atoms=1, pad_pairs=2)
    params = syn.gmnn_params(seed=31, random_bias=True, random_scale_shift=True)
    em = params["params"]["energy_model"]
    tparams = to_t(params)
    leaves = []

    def mark(tree):
        for k, v in tree.items():
            if isinstance(v, dict):
                mark(v)
            else:
                tree[k] = v.clone().requires_grad_(True)
                leaves.append((k, tree[k]))

    mark(tparams)
    m = build_model(np.zeros(3), scale=em["scale_shift"]["scale_per_element"], shift=em["scale_shift"]["shift_per_element"])
    jax_shim.HIGHER_ORDER = True
    try:
        E, F = [], []
        for s in range(n_struct):
            res = m.apply(tparams, torch.as_tensor(inputs["positions"][s]), torch.as_tensor(inputs["numbers"][s].astype(np.int64)),
                          torch.as_tensor(inputs["idx"][s].astype(np.int64)), torch.as_tensor(inputs["box"][s]),
                          torch.as_tensor(inputs["offsets"][s]))
            E.append(res["energy"])
            F.append(res["forces"])
        predictions = {"energy": torch.stack(E), "forces": torch.stack(F)}
        loss_fn = LossCollection([Loss(name="energy", loss_type="mse", weight=1.0, atoms_exponent=1),
                                  Loss(name="forces", loss_type="mse", weight=4.0, atoms_exponent=1)])
        tin = {"n_atoms": torch.as_tensor(inputs["n_atoms"].astype(np.float64))}
        tlab = {"energy": torch.as_tensor(labels["energy"]), "forces": torch.as_tensor(labels["forces"])}
        loss = loss_fn(tin, predictions, tlab)
        grads = torch.autograd.grad(loss, [t for _, t in leaves], allow_unused=True)
    finally:
        jax_shim.HIGHER_ORDER = False
    names = {"atomic_type_embedding": "emb", "scale_per_element": "scale", "shift_per_element": "shift"}
    out = {}
    li = {"w": 0, "b": 0}
    for (k, t), g in zip(leaves, grads):
        if k in ("w", "b"):
            name = f"{k}{li[k]}"
            li[k] += 1
        else:
            name = names[k]
        out["grad_" + name] = np.zeros(tuple(t.shape)) if g is None else g.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "train_step.npz"), n_struct=n_struct, seed=seed, param_seed=31, loss=float(loss.detach()),
             energy=predictions["energy"].detach().numpy(), forces=predictions["forces"].detach().numpy(), **out)


def case_basis_variants(n_mol=6, seed=17):
    """E+F with the reference's own BesselBasis (n_basis=16, r_max=5: the config-file default, model_config.py:182) and
    GaussianBasis(spacing="exponential") on a small multi-image water cell; also the raw basis values on a distance grid."""
    from apax.layers.descriptor.basis_functions import BesselBasis

    pos, Z, cell = syn.water_box(n_mol, seed)
    out = {"positions": pos, "Z": Z, "cell": cell}
    dr_grid = np.linspace(0.3, 5.9, 29).astype(np.float32)
    for name, basis, nb, r_max in (("bessel", BesselBasis(n_basis=16, r_max=5.0, dtype="fp32"), 16, 5.0),
                                   ("gexp", GaussianBasis(n_basis=7, r_min=0.5, r_max=6.0, dtype="fp32", spacing="exponential"), 7, 6.0)):
        idx, offsets = no.vesin_idx_offsets(pos, cell, r_max)
        frac = np.asarray(space.transform(torch.as_tensor(np.linalg.inv(cell)), torch.as_tensor(pos)))
        params = syn.gmnn_params(seed=41, n_basis=nb, random_bias=True, random_scale_shift=True)
        em = params["params"]["energy_model"]
        descriptor = GaussianMomentDescriptor(
            radial_fn=RadialFunction(n_radial=5, basis_fn=basis, n_species=119, emb_init="uniform", use_embed_norm=True, dtype="fp32"),
            n_contr=8, dtype="fp32", apply_mask=True)
        readout = AtomisticReadout(units=[256, 256], activation_fn=variance_preserving_swish, w_init="lecun", b_init="zeros",
                                   use_ntk=False, n_shallow_ensemble=0, dtype="fp32")
        scale_shift = PerElementScaleShift(n_species=119, scale=em["scale_shift"]["scale_per_element"],
                                           shift=em["scale_shift"]["shift_per_element"], dtype="fp64")
        m = EnergyDerivativeModel(EnergyModel(representation=descriptor, readout=readout, scale_shift=scale_shift,
                                              init_box=np.array(cell.T), inference_disp_fn=None, mask_atoms=True))
        res = np_out(m.apply(to_t(params), torch.as_tensor(frac), torch.as_tensor(Z.astype(np.int64)),
                             torch.as_tensor(idx.astype(np.int64)), torch.as_tensor(cell), torch.as_tensor(offsets)))
        vals = basis.apply({"params": {}}, torch.as_tensor(dr_grid))
        out.update({f"{name}_idx": idx, f"{name}_offsets": offsets, f"{name}_frac": frac, f"{name}_energy": res["energy"],
                    f"{name}_forces": res["forces"], f"{name}_basis": np.asarray(vals.detach() if torch.is_tensor(vals) else vals)})
    np.savez(os.path.join(HERE, "basis_variants.npz"), dr_grid=dr_grid, param_seed=41, **out)


def case_stress(n_mol=64, seed=5):
    """stress_times_vol (apax/layers/properties.py:11-42) through EnergyDerivativeModel(calc_stress=True) (models.py:160-169) on
    the min-image path: perturbation applied as (I + eps) . dR by periodic_general (space.py:260-282)."""
    pos, Z, cell = syn.water_box(n_mol, seed)
    box = torch.as_tensor(cell.T.copy())
    disp_fn, _ = space.periodic_general(box, fractional_coordinates=True)
    nfn = partition.neighbor_list(disp_fn, box, 6.0, 0.0, fractional_coordinates=True, disable_cell_list=True,
                                  format=partition.Sparse)
    frac = space.transform(torch.linalg.inv(box), torch.as_tensor(pos))
    nbrs = nfn.allocate(frac, box=box)
    params = syn.gmnn_params(seed=1234, random_bias=True)
    m = build_model(np.array(cell.T), inference_disp_fn=disp_fn)
    m.calc_stress = True
    offsets = torch.zeros((nbrs.idx.shape[1], 3), dtype=torch.int64)
    res = np_out(m.apply(to_t(params), frac, torch.as_tensor(Z.astype(np.int64)), nbrs.idx.long(), box, offsets))
    np.savez(os.path.join(HERE, "water_stress.npz"), frac=frac.numpy(), Z=Z, cell=cell, box=box.numpy(), idx=nbrs.idx.numpy(),
             param_seed=1234, energy=res["energy"], forces=res["forces"], stress=res["stress"])


def case_stress_offsets(n_mol=32, seed=7):
    """stress_times_vol through EnergyDerivativeModel(calc_stress=True) on the vesin / offsets path (inference_disp_fn=None): the
    reference's disp_fn applies the perturbation as dR + (I + eps) . dR and adds the offsets afterwards
    (apax/layers/distances.py:12-22, 62-66), so dU/d(eps) is taken at doubled displacements -- recorded as the source computes it."""
    pos, Z, cell = syn.water_box(n_mol, seed)
    idx, offsets = no.vesin_idx_offsets(pos, cell, 6.0)
    frac = np.asarray(space.transform(torch.as_tensor(np.linalg.inv(cell)), torch.as_tensor(pos)))
    params = syn.gmnn_params(seed=77, random_bias=True)
    m = build_model(np.array(cell.T), inference_disp_fn=None)
    m.calc_stress = True
    res = np_out(m.apply(to_t(params), torch.as_tensor(frac), torch.as_tensor(Z.astype(np.int64)), torch.as_tensor(idx.astype(np.int64)),
                         torch.as_tensor(cell), torch.as_tensor(offsets)))
    np.savez(os.path.join(HERE, "water_stress_offsets.npz"), frac=frac, Z=Z, cell=cell, idx=idx, offsets=offsets, param_seed=77,
             energy=res["energy"], forces=res["forces"], stress=res["stress"])


def case_corrections(n_mol=8, seed=23):
    """EnergyDerivativeModel with the reference's own ZBLRepulsion + ExponentialRepulsion (apax/layers/empirical.py) in
    EnergyModel.corrections, on a small multi-image water cell squeezed so that O-H pairs sit inside the corrections' range."""
    from apax.layers.empirical import ExponentialRepulsion, ZBLRepulsion

    pos, Z, cell = syn.water_box(n_mol, seed)
    idx, offsets = no.vesin_idx_offsets(pos, cell, 6.0)
    frac = np.asarray(space.transform(torch.as_tensor(np.linalg.inv(cell)), torch.as_tensor(pos)))
    rng = np.random.default_rng(3)
    params = syn.gmnn_params(seed=43, random_bias=True, random_scale_shift=True)
    em = params["params"]["energy_model"]
    em["corrections_0"] = {"coefficients": rng.normal(size=(4, 1)).astype(np.float32), "rep_scale": np.array([-1.7], dtype=np.float32)}
    em["corrections_1"] = {"rep_scale": rng.uniform(0.3, 1.2, size=119).astype(np.float32) * rng.choice([-1, 1], size=119),
                           "rep_prefactor": rng.uniform(1.0, 20.0, size=119).astype(np.float32)}
    em["corrections_1"]["rep_scale"] = em["corrections_1"]["rep_scale"].astype(np.float32)
    descriptor = GaussianMomentDescriptor(
        radial_fn=RadialFunction(n_radial=5, basis_fn=GaussianBasis(n_basis=7, r_min=0.5, r_max=6.0, dtype="fp32"),
                                 n_species=119, emb_init="uniform", use_embed_norm=True, dtype="fp32"),
        n_contr=8, dtype="fp32", apply_mask=True)
    readout = AtomisticReadout(units=[256, 256], activation_fn=variance_preserving_swish, w_init="lecun", b_init="zeros",
                               use_ntk=False, n_shallow_ensemble=0, dtype="fp32")
    scale_shift = PerElementScaleShift(n_species=119, scale=em["scale_shift"]["scale_per_element"],
                                       shift=em["scale_shift"]["shift_per_element"], dtype="fp64")
    corrections = [ZBLRepulsion(dtype="fp32", apply_mask=True, r_max=1.8), ExponentialRepulsion(dtype="fp32", apply_mask=True, r_max=1.5)]
    m = EnergyDerivativeModel(EnergyModel(representation=descriptor, readout=readout, scale_shift=scale_shift, corrections=corrections,
                                          init_box=np.array(cell.T), inference_disp_fn=None, mask_atoms=True))
    res = np_out(m.apply(to_t(params), torch.as_tensor(frac), torch.as_tensor(Z.astype(np.int64)),
                         torch.as_tensor(idx.astype(np.int64)), torch.as_tensor(cell), torch.as_tensor(offsets)))
    np.savez(os.path.join(HERE, "water_corrections.npz"), frac=frac, Z=Z, cell=cell, idx=idx, offsets=offsets, param_seed=43,
             zbl_coefficients=em["corrections_0"]["coefficients"], zbl_rep_scale=em["corrections_0"]["rep_scale"],
             exp_rep_scale=em["corrections_1"]["rep_scale"], exp_rep_prefactor=em["corrections_1"]["rep_prefactor"],
             zbl_r_max=1.8, exp_r_max=1.5, energy=res["energy"], forces=res["forces"])


if __name__ == "__main__":
    torch.manual_seed(0)
    case_descriptor_unit()
    print("descriptor_unit ok")
    case_gas_padded()
    print("gas_padded ok")
    case_multi_image()
    print("multi_image ok")
    case_min_image()
    print("min_image ok")
    case_ensemble()
    print("ensemble ok")
    case_md_nvt()
    print("md_nvt ok")
    case_train_step()
    print("train_step ok")
    case_basis_variants()
    print("basis_variants ok")
    case_stress()
    print("stress ok")
    case_corrections()
    print("corrections ok")
    case_stress_offsets()
    print("stress_offsets ok")
