"""Training step (BASELINE configs[1]) on the GPU against the training oracle and the fixture written by the
reference's own loss / model source (tests/golden/train_step.npz): loss, predictions, every parameter gradient
(second order through the forces), and the Adam update.

Tolerances: the reference computes this path in fp32 (descriptor, MLP) with fp64 scale/shift and loss; two fp32
evaluations with different summation orders agree to a few 1e-6 of a tensor's largest gradient entry
(the oracle itself is 1e-6 away from the reference source on the fixture).  Bound used: |dg| <= 2e-5 * max|g| per
tensor, and the error against the all-fp64 oracle must not exceed 2x the fp32 oracle's own.
"""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

REL = 2e-5
NAMES = ("emb", "w0", "b0", "w1", "b1", "w2", "b2", "scale", "shift")


def _tree_to_flat(tree):
    em = tree["params"]["energy_model"]
    out = {"emb": em["representation"]["radial_fn"]["atomic_type_embedding"], "scale": em["scale_shift"]["scale_per_element"],
           "shift": em["scale_shift"]["shift_per_element"]}
    for i in range(3):
        out[f"w{i}"] = em["readout"][f"dense_{i}"]["w"]
        out[f"b{i}"] = em["readout"][f"dense_{i}"]["b"]
    return {k: np.asarray(v) for k, v in out.items()}


def _gpu_loss_grads(params, inputs, labels, loss_terms, world=1):
    from apax_b200.optimizer.get_optimizer import get_opt
    from apax_b200.train.loss import Loss, LossCollection
    from apax_b200.train.trainer import TrainState, TrainStep, flatten_batch

    loss_fn = LossCollection([Loss(t.name, "mse", t.weight, t.atoms_exponent) for t in loss_terms])
    state = TrainState(params, get_opt(params, 10, 100, emb_lr=0.0, nn_lr=0.0, scale_lr=0.0, shift_lr=0.0))
    pos, idx = np.asarray(inputs["positions"]), np.asarray(inputs["idx"])
    step = TrainStep(state, loss_fn, pos.shape[0], pos.shape[1], idx.shape[2], world_size=world)
    loss, pred = step(inputs, labels)      # all learning rates 0: parameters stay put, gradients are left in the state
    torch.cuda.synchronize()
    return float(loss.item()), pred["energy"].cpu().numpy(), pred["forces"].cpu().numpy(), _tree_to_flat(state.grads), state


def _check_grads(got, ref32, ref64=None, where=""):
    for k in NAMES:
        g, r = got[k].reshape(-1), ref32[k].reshape(-1)
        scale = np.abs(r).max() + 1e-30
        err = np.abs(g - r).max() / scale
        assert err <= REL, f"{where} grad {k}: {err:.2e} of max|g|={scale:.3e}"
        if ref64 is not None:
            t = ref64[k].reshape(-1)
            e_gpu, e_ref = np.abs(g - t).max(), np.abs(r - t).max()
            assert e_gpu <= 2.0 * e_ref + 2e-6 * scale, f"{where} grad {k}: err vs fp64 {e_gpu:.2e}, oracle's own {e_ref:.2e}"


def test_golden_reference_source(golden):
    """Same batch as tests/golden/make_golden.py::case_train_step (reference Loss + EnergyDerivativeModel source)."""
    from oracle import train_oracle as tor

    d = golden("train_step")
    inputs, labels = tor.ethanol_training_batch(int(d["n_struct"]), seed=int(d["seed"]), pad_atoms=1, pad_pairs=2)
    params = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True, random_scale_shift=True)
    loss, e, f, grads, _ = _gpu_loss_grads(params, inputs, labels, tor.DEFAULT_LOSS)
    assert abs(loss - float(d["loss"])) <= 1e-6 * abs(float(d["loss"]))
    assert np.abs(e - d["energy"]).max() <= 1e-6 * np.abs(d["energy"]).max()
    assert np.abs(f - d["forces"]).max() <= 1e-5 * np.abs(d["forces"]).max() + 1e-6
    _check_grads(grads, {k: d["grad_" + k] for k in NAMES}, where="golden")


def test_ethanol_batch_vs_oracle_fp32_and_fp64():
    from oracle import gmnn_oracle as go
    from oracle import train_oracle as tor

    inputs, labels = tor.ethanol_training_batch(8, seed=5, pad_atoms=2, pad_pairs=5)
    inputs["n_atoms"][3] = 9
    params = syn.gmnn_params(seed=77, random_bias=True, random_scale_shift=True)
    terms = (tor.LossTerm("energy", 0.7, 1.0), tor.LossTerm("forces", 3.0, 0.5))
    ref = tor.loss_and_grads(params, inputs, labels, terms)
    ref64 = tor.loss_and_grads(params, inputs, labels, terms, cfg=go.fp64_config())
    loss, e, f, grads, _ = _gpu_loss_grads(params, inputs, labels, terms)
    assert abs(loss - ref["loss"]) <= 2e-6 * abs(ref["loss"])
    assert np.abs(e - ref["energy"]).max() <= 1e-6 * np.abs(ref["energy"]).max()
    assert np.abs(f - ref["forces"]).max() <= 1e-5 * np.abs(ref["forces"]).max() + 1e-6
    assert np.all(f[:, 9:] == 0.0)          # padding atoms: exact zeros
    _check_grads(grads, ref["grads"], ref64["grads"], where="ethanol")


def _periodic_batch(n_struct, seed):
    """Small periodic water cells with multi-image lists (the vesin path of preprocessing.py:47-54): fractional
    positions, box = cell.T, Cartesian offsets, ragged atom counts padded like input_pipeline.py:423-470."""
    from oracle import neighbor_oracle as no

    rng = np.random.default_rng(seed)
    items = []
    for s in range(n_struct):
        pos, Z, cell = syn.water_box(4 + (s % 2), seed + s)
        cell = cell + rng.normal(scale=0.2, size=(3, 3))
        idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
        items.append((pos @ np.linalg.inv(cell), Z, cell.T.copy(), idx, off))
    nmax = max(len(it[1]) for it in items)
    pmax = max(it[3].shape[1] for it in items) + 3
    inputs = {"positions": np.zeros((n_struct, nmax, 3)), "numbers": np.zeros((n_struct, nmax), dtype=np.int32),
              "idx": np.zeros((n_struct, 2, pmax), dtype=np.int32), "box": np.zeros((n_struct, 3, 3)),
              "offsets": np.zeros((n_struct, pmax, 3)), "n_atoms": np.zeros(n_struct, dtype=np.int32)}
    labels = {"energy": rng.normal(size=n_struct) * 3, "forces": np.zeros((n_struct, nmax, 3))}
    for s, (frac, Z, box, idx, off) in enumerate(items):
        n, p = len(Z), idx.shape[1]
        inputs["positions"][s, :n] = frac
        inputs["numbers"][s, :n] = Z
        inputs["idx"][s, :, :p] = idx
        inputs["box"][s] = box
        inputs["offsets"][s, :p] = off
        inputs["n_atoms"][s] = n
        labels["forces"][s, :n] = rng.normal(size=(n, 3))
    return inputs, labels


def test_periodic_cells_with_offsets_vs_oracle():
    from oracle import gmnn_oracle as go
    from oracle import train_oracle as tor

    inputs, labels = _periodic_batch(3, seed=11)
    params = syn.gmnn_params(seed=5, random_bias=True, random_scale_shift=True)
    ref = tor.loss_and_grads(params, inputs, labels, mode="offsets")
    ref64 = tor.loss_and_grads(params, inputs, labels, mode="offsets", cfg=go.fp64_config())
    loss, e, f, grads, _ = _gpu_loss_grads(params, inputs, labels, tor.DEFAULT_LOSS)
    assert abs(loss - ref["loss"]) <= 2e-6 * abs(ref["loss"])
    # these random-weight energies nearly cancel (|E| ~ 1 from ~15 atomic terms of either sign), so 1e-6 of |E| is below the
    # fp32 noise of the sum; accept either that or twice the fp32 oracle's own distance from the fp64 truth
    e_tol = max(1e-6 * np.abs(ref["energy"]).max(), 2.0 * np.abs(ref["energy"] - ref64["energy"]).max())
    assert np.abs(e - ref64["energy"]).max() <= e_tol
    _check_grads(grads, ref["grads"], ref64["grads"], where="periodic")


def test_adam_steps_follow_oracle_and_graph_replay_is_identical():
    from apax_b200.optimizer.get_optimizer import get_opt
    from apax_b200.train.loss import Loss, LossCollection
    from apax_b200.train.trainer import TrainState, TrainStep
    from oracle import train_oracle as tor

    inputs, labels = tor.ethanol_training_batch(4, seed=2)
    params = syn.gmnn_params(seed=9, random_bias=True, random_scale_shift=True)
    oc = tor.OptConfig(emb_lr=0.02, nn_lr=0.03, scale_lr=0.001, shift_lr=0.05, transition_steps=7, end_value=1e-4)
    p_ref, st_ref = params, tor.AdamState()
    losses_ref = []
    for _ in range(3):
        l, _, p_ref, st_ref = tor.train_step(p_ref, st_ref, inputs, labels, oc=oc)
        losses_ref.append(l)
    loss_fn = LossCollection([Loss("energy", "mse", 1.0, 1.0), Loss("forces", "mse", 4.0, 1.0)])

    def run(graph):
        tx = get_opt(params, 1, 7, emb_lr=0.02, nn_lr=0.03, scale_lr=0.001, shift_lr=0.05, schedule={"name": "linear", "end_value": 1e-4})
        state = TrainState(params, tx)
        step = TrainStep(state, loss_fn, 4, 9, inputs["idx"].shape[2], graph=graph)
        losses = []
        for _ in range(3):
            l, _ = step(inputs, labels)
            losses.append(float(l.item()))
        torch.cuda.synchronize()
        return losses, _tree_to_flat(state.params), state

    losses, theta, state = run(False)
    assert state.step == 3
    ref_flat = _tree_to_flat(p_ref)
    np.testing.assert_allclose(losses, losses_ref, rtol=2e-5)
    for k in NAMES:
        # the first Adam step moves every weight by ~lr regardless of the gradient's size (m/sqrt(v) = +-1), so
        # entries whose gradient is at rounding level may differ by up to 2 lr per step; compare where it is not
        d = np.abs(theta[k] - ref_flat[k])
        frac_close = np.mean(d <= 1e-4 * (np.abs(ref_flat[k]) + 1e-3))
        assert frac_close >= 0.995, f"{k}: only {frac_close:.4f} of the entries follow the oracle"
    losses_g, theta_g, _ = run(True)
    assert losses_g == losses
    for k in NAMES:
        assert np.array_equal(theta_g[k], theta[k]), k


def test_bitwise_reproducible_and_rank_partition_adds_up():
    """Two evaluations are bitwise equal; gradients of two half batches evaluated with the global batch divisor add up
    to the full-batch gradient (what the all-reduce relies on)."""
    from oracle import train_oracle as tor

    inputs, labels = tor.ethanol_training_batch(6, seed=13)
    params = syn.gmnn_params(seed=3, random_bias=True)
    l1, _, _, g1, _ = _gpu_loss_grads(params, inputs, labels, tor.DEFAULT_LOSS)
    l2, _, _, g2, _ = _gpu_loss_grads(params, inputs, labels, tor.DEFAULT_LOSS)
    assert l1 == l2
    for k in NAMES:
        assert np.array_equal(g1[k], g2[k]), k
    halves = []
    for sl in (slice(0, 3), slice(3, 6)):
        inp = {k: v[sl] for k, v in inputs.items()}
        lab = {k: v[sl] for k, v in labels.items()}
        halves.append(_gpu_loss_grads(params, inp, lab, tor.DEFAULT_LOSS, world=2))
    assert abs(halves[0][0] + halves[1][0] - l1) <= 1e-12 * abs(l1)
    for k in NAMES:
        s = halves[0][3][k] + halves[1][3][k]
        assert np.abs(s - g1[k]).max() <= 2e-6 * (np.abs(g1[k]).max() + 1e-30), k


def test_val_step_returns_loss_without_updating():
    from apax_b200.optimizer.get_optimizer import get_opt
    from apax_b200.train.loss import Loss, LossCollection
    from apax_b200.train.trainer import TrainState, make_step_fns
    from oracle import train_oracle as tor

    inputs, labels = tor.ethanol_training_batch(4, seed=3)
    params = syn.gmnn_params(seed=12, random_bias=True)
    ref = tor.loss_and_grads(params, inputs, labels)
    state = TrainState(params, get_opt(params, 1, 10))
    train_step, val_step = make_step_fns(LossCollection([Loss("energy", "mse", 1.0, 1.0), Loss("forces", "mse", 4.0, 1.0)]))
    before = state.theta32.clone()
    loss, _, pred = val_step(state, (inputs, labels))
    assert abs(float(loss.item()) - ref["loss"]) <= 2e-6 * abs(ref["loss"])
    assert torch.equal(before, state.theta32) and state.step == 0
    assert np.abs(pred["forces"].cpu().numpy() - ref["forces"]).max() <= 1e-5 * np.abs(ref["forces"]).max() + 1e-6
    (state, _), loss2 = train_step((state, None), (inputs, labels))
    assert state.step == 1 and not torch.equal(before, state.theta32)
    assert abs(float(loss2.item()) - ref["loss"]) <= 2e-6 * abs(ref["loss"])


def test_periodic_128_atom_cells_training_gradients():
    """Training on realistic periodic structures: two 128-atom water cells (multi-image lists, ~11.6 k entries each, more
    entries per structure than the shared-memory staging of the list grouping holds) against the oracle."""
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no
    from oracle import train_oracle as tor

    rng = np.random.default_rng(5)
    items = []
    for s in range(2):
        pos, Z, cell = syn.cell_128(40 + s)
        idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
        items.append((pos @ np.linalg.inv(cell), Z, cell.T.copy(), idx, off))
    pmax = max(it[3].shape[1] for it in items)
    assert pmax > 3072
    inputs = {"positions": np.stack([it[0] for it in items]), "numbers": np.stack([it[1] for it in items]).astype(np.int32),
              "idx": np.zeros((2, 2, pmax), dtype=np.int32), "box": np.stack([it[2] for it in items]),
              "offsets": np.zeros((2, pmax, 3)), "n_atoms": np.array([128, 128], dtype=np.int32)}
    for s, it in enumerate(items):
        inputs["idx"][s, :, : it[3].shape[1]] = it[3]
        inputs["offsets"][s, : it[3].shape[1]] = it[4]
    labels = {"energy": rng.normal(size=2) * 10, "forces": rng.normal(size=(2, 128, 3))}
    params = syn.gmnn_params(seed=6, random_bias=True, random_scale_shift=True)
    ref = tor.loss_and_grads(params, inputs, labels, mode="offsets")
    ref64 = tor.loss_and_grads(params, inputs, labels, mode="offsets", cfg=go.fp64_config())
    loss, e, f, grads, _ = _gpu_loss_grads(params, inputs, labels, tor.DEFAULT_LOSS)
    assert abs(loss - ref["loss"]) <= 2e-6 * abs(ref["loss"])
    assert np.abs(f - ref64["forces"]).max() <= 2.0 * np.abs(ref["forces"] - ref64["forces"]).max() + 1e-5 * np.abs(ref["forces"]).max()
    _check_grads(grads, ref["grads"], ref64["grads"], where="periodic-128")
