"""Training oracle (oracle/train_oracle.py) on CPU: pinned against the fixture written by the reference's own
loss.py / models.py source, against fp64 finite differences, and (Adam) against torch.optim.Adam; host logic of
the training mirrors (batch flattening, loss / optimizer configuration records)."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn
from oracle import gmnn_oracle as go
from oracle import train_oracle as tor

NAMES = ("emb", "w0", "b0", "w1", "b1", "w2", "b2", "scale", "shift")


def test_oracle_matches_reference_source_fixture(golden):
    d = golden("train_step")
    inputs, labels = tor.ethanol_training_batch(int(d["n_struct"]), seed=int(d["seed"]), pad_atoms=1, pad_pairs=2)
    params = syn.gmnn_params(seed=int(d["param_seed"]), random_bias=True, random_scale_shift=True)
    r = tor.loss_and_grads(params, inputs, labels)
    assert abs(r["loss"] - float(d["loss"])) <= 1e-7 * abs(float(d["loss"]))
    assert np.abs(r["energy"] - d["energy"]).max() <= 1e-6
    assert np.abs(r["forces"] - d["forces"]).max() <= 1e-6
    for k in NAMES:
        g = d["grad_" + k]
        assert np.abs(r["grads"][k] - g).max() <= 2e-6 * np.abs(g).max(), k
    # only the species pairs present in ethanol receive an embedding gradient
    nz = np.argwhere(np.abs(r["grads"]["emb"]).sum(axis=(2, 3)) > 0)
    assert {tuple(x) for x in nz} == {(a, b) for a in (1, 6, 8) for b in (1, 6, 8)} - {(8, 8)}


def test_fp64_oracle_gradient_vs_finite_differences():
    inputs, labels = tor.ethanol_training_batch(2, seed=4)
    params = syn.gmnn_params(seed=2, random_bias=True, random_scale_shift=True)
    cfg = go.fp64_config()
    p64 = tor.flat_to_params({k: v.astype(np.float64) for k, v in tor.params_to_flat(params).items()})
    r = tor.loss_and_grads(p64, inputs, labels, cfg=cfg)
    rng = np.random.default_rng(0)
    flat = tor.params_to_flat(p64)
    for k, index in (("w0", (17, 33)), ("w1", (5, 200)), ("b1", (7,)), ("w2", (100, 0)), ("emb", (6, 1, 2, 3)),
                     ("emb", (1, 8, 4, 0)), ("scale", (6, 0)), ("shift", (1, 0))):
        h = 1e-5
        vals = []
        for sgn in (+1, -1):
            f2 = {n: v.copy() for n, v in flat.items()}
            f2[k][index] += sgn * h
            vals.append(tor.loss_and_grads(tor.flat_to_params(f2), inputs, labels, cfg=cfg)["loss"])
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - r["grads"][k][index]) <= 1e-6 * max(1.0, abs(fd)), (k, index, fd, r["grads"][k][index])


def test_adam_update_matches_torch_adam_with_constant_lr():
    rng = np.random.default_rng(1)
    theta = {"w0": rng.normal(size=(5, 4)).astype(np.float32), "emb": rng.normal(size=(3, 2)).astype(np.float32),
             "scale": rng.normal(size=(4, 1)), "shift": rng.normal(size=(4, 1))}
    oc = tor.OptConfig(emb_lr=0.02, nn_lr=0.03, scale_lr=0.001, shift_lr=0.05, end_value=0.0, transition_steps=10 ** 9)
    tparams = {k: torch.nn.Parameter(torch.as_tensor(v.copy())) for k, v in theta.items()}
    opts = {k: torch.optim.Adam([tparams[k]], lr=oc.group_lr(k.rstrip("0123456789")), betas=(0.9, 0.999), eps=1e-8) for k in theta}
    state = tor.AdamState()
    for it in range(4):
        grads = {k: (rng.normal(size=v.shape) * (2000.0 if it == 2 else 1.0)).astype(v.dtype) for k, v in theta.items()}
        theta, state = tor.adam_update(theta, grads, state, oc)
        for k in tparams:
            tparams[k].grad = torch.as_tensor(np.clip(grads[k], -1000.0, 1000.0))
            opts[k].step()
    for k in theta:
        tol = 2e-6 if theta[k].dtype == np.float32 else 1e-9   # the (huge) transition_steps leaves lr constant to 1e-9
        np.testing.assert_allclose(theta[k], tparams[k].detach().numpy(), rtol=tol, atol=tol)
    assert state.count == 4


def test_linear_schedule_and_frozen_groups():
    oc = tor.OptConfig(transition_steps=100, end_value=1e-6, transition_begin=10)
    assert tor.linear_schedule(0, 0.01, oc) == pytest.approx(0.01)
    assert tor.linear_schedule(10, 0.01, oc) == pytest.approx(0.01)
    assert tor.linear_schedule(60, 0.01, oc) == pytest.approx((0.01 - 1e-6) * 0.5 + 1e-6)
    assert tor.linear_schedule(10 ** 6, 0.01, oc) == pytest.approx(1e-6)
    theta = {"w0": np.ones((2, 2), dtype=np.float32), "shift": np.ones((2, 1))}
    new, _ = tor.adam_update(theta, {k: np.ones_like(v) for k, v in theta.items()}, tor.AdamState(),
                             tor.OptConfig(nn_lr=0.0, shift_lr=0.05))
    assert np.array_equal(new["w0"], theta["w0"]) and not np.array_equal(new["shift"], theta["shift"])


def test_flatten_batch_and_config_records():
    from apax_b200.optimizer.get_optimizer import get_opt
    from apax_b200.train.loss import Loss, LossCollection
    from apax_b200.train.trainer import flatten_batch, shard_batch

    inputs, labels = tor.ethanol_training_batch(4, seed=0, pad_atoms=1, pad_pairs=2)
    flat = flatten_batch(inputs, labels)
    B, nmax, pmax = 4, 10, inputs["idx"].shape[2]
    assert flat["R"].shape == (B * nmax, 3) and flat["idx"].shape == (2, B * pmax)
    assert list(flat["struct_ptr"]) == [0, 10, 20, 30, 40] and flat["pair_ptr"][-1] == B * pmax
    s = 2
    assert np.array_equal(flat["idx"][:, s * pmax:(s + 1) * pmax], inputs["idx"][s] + s * nmax)
    assert np.array_equal(flat["R"][s * nmax:(s + 1) * nmax], inputs["positions"][s])
    # periodic structures: fractional -> Cartesian with box . frac
    inp2 = {k: v.copy() for k, v in inputs.items()}
    box = np.array([[4.0, 0.5, 0.0], [0.0, 5.0, 0.3], [0.0, 0.0, 6.0]])
    inp2["box"][1] = box
    f2 = flatten_batch(inp2, labels)
    assert np.allclose(f2["R"][nmax:2 * nmax], inputs["positions"][1] @ box.T)
    assert np.array_equal(f2["R"][:nmax], inputs["positions"][0])
    i0, l0 = shard_batch(inputs, labels, 1, 2)
    assert np.array_equal(i0["positions"], inputs["positions"][2:4]) and np.array_equal(l0["energy"], labels["energy"][2:4])
    with pytest.raises(ValueError):
        shard_batch(inputs, labels, 0, 3)
    c = LossCollection([Loss("energy", "mse", 1.0, 1), Loss("forces", "mse", 4.0, 0.5)]).to_c()
    assert (c.energy_weight, c.energy_atoms_exponent, c.forces_weight, c.forces_atoms_exponent) == (1.0, 1.0, 4.0, 0.5)
    with pytest.raises(NotImplementedError):
        Loss("energy", "huber")
    with pytest.raises(NotImplementedError):
        Loss("stress", "mse")
    tx = get_opt(None, 10, 7, emb_lr=0.001, nn_lr=0.002, schedule={"name": "linear", "transition_begin": 3, "end_value": 1e-5})
    oc = tx.to_c()
    assert oc.transition_steps == 70 and oc.transition_begin == 3 and oc.end_value == 1e-5 and oc.nn_lr == 0.002
    with pytest.raises(NotImplementedError):
        get_opt(None, 1, 1, name="sam")
