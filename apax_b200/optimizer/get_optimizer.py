"""Mirror of apax/optimizer/get_optimizer.py (`get_opt`, :90-176) for the GMNN parameter tree.

The reference builds an optax `multi_transform` of per-group `chain(clip, adam(schedule), zero_nans)`; the CUDA
path applies exactly that update in one kernel (`apax_train_adam_step`, csrc/train.cu), so `get_opt` returns the
plain configuration record the kernel takes.  Only `name="adam"` with the linear schedule (lr_config.py:10-26) is
on the path named by BASELINE configs[1].
"""
from __future__ import annotations

import dataclasses

from .. import _lib


@dataclasses.dataclass
class AdamLinear:
    emb_lr: float
    nn_lr: float
    scale_lr: float
    shift_lr: float
    gradient_clipping: float
    transition_steps: int
    transition_begin: int = 0
    end_value: float = 1e-6
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8

    def to_c(self) -> _lib.TrainOptConfig:
        return _lib.TrainOptConfig(self.emb_lr, self.nn_lr, self.scale_lr, self.shift_lr, self.gradient_clipping,
                                   self.end_value, int(self.transition_steps), int(self.transition_begin), self.b1, self.b2,
                                   self.eps)


def get_opt(params, n_epochs: int, steps_per_epoch: int, emb_lr: float = 0.02, nn_lr: float = 0.03,
            scale_lr: float = 0.001, shift_lr: float = 0.05, zbl_lr: float = 0.001, rep_scale_lr: float = 0.001,
            rep_prefactor_lr: float = 0.0001, gradient_clipping=1000.0, freeze_layers: list = [], name: str = "adam",
            kwargs: dict = {}, schedule: dict = {}) -> AdamLinear:
    """Same signature and defaults as the reference (get_optimizer.py:90-106)."""
    if name != "adam":
        raise NotImplementedError(f"optimizer '{name}': only optax.adam is on the CUDA training path")
    if freeze_layers:
        raise NotImplementedError("freeze_layers: set the group's learning rate to 0 instead (get_optimizer.py:74-75)")
    sched = dict(schedule) or {"name": "linear"}
    if sched.pop("name", "linear") != "linear":
        raise KeyError("unknown learning rate schedule on the CUDA training path (only 'linear')")
    extra = set(kwargs) - {"b1", "b2", "eps"}
    if extra:
        raise NotImplementedError(f"adam kwargs {sorted(extra)}")
    return AdamLinear(emb_lr=emb_lr, nn_lr=nn_lr, scale_lr=scale_lr, shift_lr=shift_lr, gradient_clipping=gradient_clipping,
                      transition_steps=n_epochs * steps_per_epoch, transition_begin=int(sched.get("transition_begin", 0)),
                      end_value=float(sched.get("end_value", 1e-6)), **{k: kwargs[k] for k in kwargs})
