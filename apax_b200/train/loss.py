"""Mirror of apax/train/loss.py (`Loss`, `LossCollection`, :199-259) for the GMNN training step.

The reference evaluates these objects inside the jitted step; here they only carry the configuration that
`apax_train_loss_and_grad` (csrc/train.cu) needs: loss type "mse" (loss.py:12-23) on "energy" and "forces",
weight and atoms_exponent.  Other loss types / targets are outside the hot path and raise, like the
reference does for unknown types (loss.py:211-215).
"""
from __future__ import annotations

import dataclasses
from typing import List

from .. import _lib

SUPPORTED_TYPES = ("mse",)
SUPPORTED_NAMES = ("energy", "forces")


@dataclasses.dataclass
class Loss:
    name: str
    loss_type: str
    weight: float = 1.0
    atoms_exponent: float = 1.0
    parameters: dict = dataclasses.field(default_factory=lambda: {})

    def __post_init__(self):
        if self.loss_type not in SUPPORTED_TYPES:
            raise NotImplementedError(f"the loss function '{self.loss_type}' is not known to the CUDA training path.")
        if self.name not in SUPPORTED_NAMES:
            raise NotImplementedError(f"loss target '{self.name}' is not on the GMNN energy+forces path.")


@dataclasses.dataclass
class LossCollection:
    loss_list: List[Loss]

    def to_c(self) -> _lib.TrainLossConfig:
        """Sum of the terms per target (LossCollection.__call__ adds them up, loss.py:251-259); several terms on the
        same target are only expressible when their atoms_exponent agrees."""
        cfg = {"energy": [0.0, 1.0], "forces": [0.0, 1.0]}
        seen = {}
        for term in self.loss_list:
            if term.name in seen and seen[term.name] != term.atoms_exponent:
                raise NotImplementedError("two loss terms on one target with different atoms_exponent")
            seen[term.name] = term.atoms_exponent
            cfg[term.name][0] += float(term.weight)
            cfg[term.name][1] = float(term.atoms_exponent)
        return _lib.TrainLossConfig(cfg["energy"][0], cfg["energy"][1], cfg["forces"][0], cfg["forces"][1])
