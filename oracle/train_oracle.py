"""CPU ORACLE (test infrastructure, NOT product code) for apax's GMNN *training step*
(BASELINE configs[1]): loss of the vmapped EnergyDerivativeModel,
its gradient w.r.t. every parameter (second order: the force loss differentiates
F = -dE/dR w.r.t. the weights), and the optax-Adam update with per-group learning rates.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import this file; `apax_b200` never does.

Reference code restated (paths relative to /root/reference):
  calc_loss                     apax/train/trainer.py:253-263
  batched model                 apax/train/run.py:204  (jax.vmap(model.apply, (None,0,0,0,0,0)))
  Loss.__call__ / mse           apax/train/loss.py:12-23, 199-244
  update_step                   apax/train/trainer.py:332-339 (value_and_grad -> apply_gradients)
  get_opt / OptimizerFactory    apax/optimizer/get_optimizer.py:62-176
  LinearLR                      apax/config/lr_config.py:10-26

Pinning status: the model + loss part is pinned by `tests/golden/train_step.npz`, written by
`tests/golden/make_golden.py` from the reference's OWN `apax/train/loss.py` and `apax/nn/models.py`
run under the torch-backed shim (autograd with create_graph for the inner jax.value_and_grad).
optax (0.2.x, `uv.lock`) is absent from /root/reference: Adam, `optax.clip`, `linear_schedule`,
`zero_nans` and `multi_transform` are restated from optax's published algorithm — *parity
unpinned* for the optimizer update (no golden vector exists upstream either).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import gmnn_oracle as go


@dataclass
class LossTerm:
    """apax/train/loss.py:199-244 with loss_type "mse" (the config default, train_config.py:297)."""

    name: str            # "energy" or "forces"
    weight: float = 1.0
    atoms_exponent: float = 1.0


DEFAULT_LOSS = (LossTerm("energy", 1.0, 1.0), LossTerm("forces", 4.0, 1.0))


def _param_tensors(params: dict, requires_grad: bool):
    em = params["params"]["energy_model"] if "params" in params else params
    out = {
        "emb": torch.as_tensor(np.asarray(em["representation"]["radial_fn"]["atomic_type_embedding"])).clone(),
        "scale": torch.as_tensor(np.asarray(em["scale_shift"]["scale_per_element"], dtype=np.float64)).reshape(-1, 1).clone(),
        "shift": torch.as_tensor(np.asarray(em["scale_shift"]["shift_per_element"], dtype=np.float64)).reshape(-1, 1).clone(),
    }
    ro = em["readout"]
    for li in range(len(ro)):
        out[f"w{li}"] = torch.as_tensor(np.asarray(ro[f"dense_{li}"]["w"])).clone()
        out[f"b{li}"] = torch.as_tensor(np.asarray(ro[f"dense_{li}"]["b"])).clone()
    if requires_grad:
        for v in out.values():
            v.requires_grad_(True)
    return out


def _tree_from_tensors(t: Dict[str, torch.Tensor]) -> dict:
    n_layers = sum(1 for k in t if k.startswith("w"))
    return {
        "representation": {"radial_fn": {"atomic_type_embedding": t["emb"]}},
        "readout": {f"dense_{i}": {"w": t[f"w{i}"], "b": t[f"b{i}"]} for i in range(n_layers)},
        "scale_shift": {"scale_per_element": t["scale"], "shift_per_element": t["shift"]},
    }


def _energy_one(tens, R, Z, idx, box, offsets, mode, cfg):
    """EnergyModel.__call__ on torch parameter tensors (keeps the autograd graph to the weights)."""
    dr_vec = go.compute_distances(R, idx, box, offsets, mode)
    g, _ = go.gaussian_moment_descriptor(dr_vec, Z, idx, tens["emb"], cfg)
    dense = [{"w": tens[f"w{i}"], "b": tens[f"b{i}"]} for i in range(sum(1 for k in tens if k.startswith("w")))]
    h = go.atomistic_readout(g, dense, cfg)
    E_i = go.per_element_scale_shift(h, Z, tens["scale"], tens["shift"], cfg)
    if cfg.mask_atoms:
        E_i = E_i * (Z != 0).to(E_i.dtype)[:, None]
    return torch.sum(E_i.to(torch.float64))


def loss_and_grads(params: dict, inputs: dict, labels: dict, loss_terms=DEFAULT_LOSS, mode: str = "gas",
                   cfg: Optional[go.GMNNConfig] = None, batch_divisor: Optional[int] = None):
    """calc_loss + jax.value_and_grad w.r.t. params (trainer.py:253-263, 332-336).

    inputs: positions [B,Nmax,3], numbers [B,Nmax], idx [B,2,Pmax], box [B,3,3], offsets [B,Pmax,3],
            n_atoms [B]  (the padded batch of apax/data/input_pipeline.py:423-470)
    labels: energy [B], forces [B,Nmax,3]
    The vmapped model is evaluated structure by structure (same values as jax.vmap).
    `batch_divisor` (default B) is the batch size of the `jnp.mean(arg, axis=0)` at loss.py:241; under
    data parallelism it is the GLOBAL batch size so that per-rank gradients simply add up.
    Returns dict(loss, energy [B], forces [B,Nmax,3], grads {name: ndarray}).
    """
    cfg = cfg or go.GMNNConfig()
    tens = _param_tensors(params, requires_grad=True)
    B = int(np.asarray(inputs["positions"]).shape[0])
    div_b = float(batch_divisor or B)
    E_pred, F_pred = [], []
    for s in range(B):
        R = torch.as_tensor(np.asarray(inputs["positions"][s], dtype=np.float64)).clone().requires_grad_(True)
        Z = torch.as_tensor(np.asarray(inputs["numbers"][s]).astype(np.int64))
        idx = torch.as_tensor(np.asarray(inputs["idx"][s]).astype(np.int64))
        box = torch.as_tensor(np.asarray(inputs["box"][s], dtype=np.float64))
        off = torch.as_tensor(np.asarray(inputs["offsets"][s], dtype=np.float64))
        E = _energy_one(tens, R, Z, idx, box, off, mode, cfg)
        (gR,) = torch.autograd.grad(E, R, create_graph=True)       # models.py:154-156
        E_pred.append(E)
        F_pred.append(-gR)
    E_pred = torch.stack(E_pred)                                    # [B] f64
    F_pred = torch.stack(F_pred)                                    # [B,Nmax,3] f64
    n_atoms = torch.as_tensor(np.asarray(inputs["n_atoms"]).astype(np.float64))
    E_lab = torch.as_tensor(np.asarray(labels["energy"], dtype=np.float64))
    F_lab = torch.as_tensor(np.asarray(labels["forces"], dtype=np.float64))
    total = torch.zeros((), dtype=torch.float64)
    for term in loss_terms:
        divisor = n_atoms ** term.atoms_exponent                    # loss.py:233
        if term.name == "energy":
            arg = (E_lab - E_pred) ** 2 / divisor
        elif term.name == "forces":
            arg = (F_lab - F_pred) ** 2 / divisor[:, None, None]
        else:
            raise NotImplementedError(term.name)
        total = total + term.weight * torch.sum(torch.sum(arg, dim=0) / div_b)   # loss.py:241-242
    names = list(tens.keys())
    grads = torch.autograd.grad(total, [tens[k] for k in names], allow_unused=True)
    gd = {k: (np.zeros(tuple(tens[k].shape), dtype=tens[k].detach().numpy().dtype) if g is None else g.detach().numpy())
          for k, g in zip(names, grads)}
    return {"loss": float(total.detach()), "energy": E_pred.detach().numpy(), "forces": F_pred.detach().numpy(), "grads": gd}


# --------------------------------------------------------------------------------------- optimizer
@dataclass
class OptConfig:
    """get_opt arguments (get_optimizer.py:90-106) with the config defaults of
    apax/config/train_config.py:238-252 and LinearLR (lr_config.py:24-26)."""

    emb_lr: float = 0.001
    nn_lr: float = 0.001
    scale_lr: float = 0.0001
    shift_lr: float = 0.003
    gradient_clipping: float = 1000.0
    transition_steps: int = 1000          # n_epochs * steps_per_epoch
    transition_begin: int = 0
    end_value: float = 1e-6
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8

    def group_lr(self, name: str) -> float:
        if name == "emb":
            return self.emb_lr
        if name == "scale":
            return self.scale_lr
        if name == "shift":
            return self.shift_lr
        return self.nn_lr                  # "w", "b"


def linear_schedule(count: int, init_value: float, oc: OptConfig) -> float:
    """optax.linear_schedule = polynomial_schedule(power=1)."""
    c = min(max(count - oc.transition_begin, 0), oc.transition_steps)
    frac = 1.0 - c / oc.transition_steps
    return (init_value - oc.end_value) * frac + oc.end_value


@dataclass
class AdamState:
    count: int = 0
    mu: Dict[str, np.ndarray] = field(default_factory=dict)
    nu: Dict[str, np.ndarray] = field(default_factory=dict)


def adam_update(theta: Dict[str, np.ndarray], grads: Dict[str, np.ndarray], state: AdamState, oc: OptConfig):
    """optax.chain(clip(c), adam(schedule), zero_nans()) per parameter group inside multi_transform
    (get_optimizer.py:73-85, 151-176); lr <= 1e-7 freezes the group (:74-75).  Arrays keep their dtype
    (fp32 weights / embedding, fp64 scale and shift).  Returns (new_theta, new_state)."""
    t = state.count + 1
    new_theta, mu_n, nu_n = {}, {}, {}
    for k, p in theta.items():
        dt = p.dtype
        g = np.clip(grads[k].astype(dt), -oc.gradient_clipping, oc.gradient_clipping)
        lr0 = oc.group_lr(k.rstrip("0123456789"))
        mu = state.mu.get(k, np.zeros_like(p))
        nu = state.nu.get(k, np.zeros_like(p))
        if lr0 <= 1e-7:
            new_theta[k], mu_n[k], nu_n[k] = p.copy(), mu, nu
            continue
        mu = (oc.b1 * mu + (1.0 - oc.b1) * g).astype(dt)
        nu = (oc.b2 * nu + (1.0 - oc.b2) * g * g).astype(dt)
        mu_hat = mu / dt.type(1.0 - oc.b1 ** t)
        nu_hat = nu / dt.type(1.0 - oc.b2 ** t)
        lr = linear_schedule(state.count, lr0, oc)
        upd = (-lr * (mu_hat / (np.sqrt(nu_hat) + dt.type(oc.eps)))).astype(dt)
        upd = np.where(np.isnan(upd), dt.type(0), upd)
        new_theta[k] = (p + upd).astype(dt)
        mu_n[k], nu_n[k] = mu, nu
    return new_theta, AdamState(count=t, mu=mu_n, nu=nu_n)


def params_to_flat(params: dict) -> Dict[str, np.ndarray]:
    return {k: v.detach().numpy().copy() for k, v in _param_tensors(params, requires_grad=False).items()}


def flat_to_params(theta: Dict[str, np.ndarray]) -> dict:
    n_layers = sum(1 for k in theta if k.startswith("w"))
    return {"params": {"energy_model": {
        "representation": {"radial_fn": {"atomic_type_embedding": theta["emb"]}},
        "readout": {f"dense_{i}": {"w": theta[f"w{i}"], "b": theta[f"b{i}"]} for i in range(n_layers)},
        "scale_shift": {"scale_per_element": theta["scale"], "shift_per_element": theta["shift"]},
    }}}


def train_step(params: dict, state: AdamState, inputs: dict, labels: dict, loss_terms=DEFAULT_LOSS,
               oc: Optional[OptConfig] = None, mode: str = "gas", cfg: Optional[go.GMNNConfig] = None):
    """update_step (trainer.py:332-336): returns (loss, predictions, new params, new optimizer state)."""
    oc = oc or OptConfig()
    res = loss_and_grads(params, inputs, labels, loss_terms, mode, cfg)
    theta = params_to_flat(params)
    new_theta, new_state = adam_update(theta, res["grads"], state, oc)
    return res["loss"], {"energy": res["energy"], "forces": res["forces"]}, flat_to_params(new_theta), new_state


# --------------------------------------------------------------------------------------- synthetic batches
def ethanol_training_batch(n_struct: int, seed: int = 0, r_max: float = 6.0, pad_atoms: int = 0, pad_pairs: int = 0):
    """BASELINE configs[1]: gas-phase ethanol-shaped molecules, full directed list inside r_max
    (apax/data/preprocessing.py:35-45 puts gas-phase structures in a shrink-wrapped cell, so the list is all
    ordered pairs closer than r_max, zero offsets), labels E ~ N(0,1), F ~ N(0,1).
    Optional padding atoms (Z=0 at the origin) / padding pairs (0,0) as input_pipeline.py:423-470 makes them."""
    from apax_b200 import synthetic as syn

    R, Z = syn.ethanol_batch(n_struct, seed)
    rng = np.random.default_rng(seed + 977)
    n = R.shape[1]
    idx_list = []
    for s in range(n_struct):
        d = np.linalg.norm(R[s][:, None, :] - R[s][None, :, :], axis=-1)
        i, j = np.nonzero((d < r_max) & ~np.eye(n, dtype=bool))
        idx_list.append(np.stack([i, j]).astype(np.int32))
    p_max = max(x.shape[1] for x in idx_list) + pad_pairs
    n_max = n + pad_atoms
    pos = np.zeros((n_struct, n_max, 3))
    num = np.zeros((n_struct, n_max), dtype=np.int32)
    idx = np.zeros((n_struct, 2, p_max), dtype=np.int32)
    pos[:, :n] = R
    num[:, :n] = Z
    for s, x in enumerate(idx_list):
        idx[s, :, : x.shape[1]] = x
    F = np.zeros((n_struct, n_max, 3))
    F[:, :n] = rng.normal(size=(n_struct, n, 3))
    inputs = {"positions": pos, "numbers": num, "idx": idx, "box": np.zeros((n_struct, 3, 3)),
              "offsets": np.zeros((n_struct, p_max, 3)), "n_atoms": np.full((n_struct,), n, dtype=np.int32)}
    labels = {"energy": rng.normal(size=(n_struct,)), "forces": F}
    return inputs, labels
