"""Synthetic training batches in the reference's padded layout (apax/data/input_pipeline.py:423-470); product-side
twin of oracle/train_oracle.py::ethanol_training_batch so that bench.py's GPU arm does not import the oracle."""
from __future__ import annotations

import numpy as np

from .. import synthetic as syn


def ethanol_training_batch(n_struct: int, seed: int = 0, r_max: float = 6.0):
    """BASELINE configs[1]: gas-phase ethanol-shaped molecules, full directed list inside r_max, zero offsets and box
    (apax/data/preprocessing.py:35-45), labels E ~ N(0,1), F ~ N(0,1)."""
    R, Z = syn.ethanol_batch(n_struct, seed)
    rng = np.random.default_rng(seed + 977)
    n = R.shape[1]
    lists = []
    for s in range(n_struct):
        d = np.linalg.norm(R[s][:, None, :] - R[s][None, :, :], axis=-1)
        i, j = np.nonzero((d < r_max) & ~np.eye(n, dtype=bool))
        lists.append(np.stack([i, j]).astype(np.int32))
    p_max = max(x.shape[1] for x in lists)
    idx = np.zeros((n_struct, 2, p_max), dtype=np.int32)
    for s, x in enumerate(lists):
        idx[s, :, : x.shape[1]] = x
    inputs = {"positions": R, "numbers": Z.astype(np.int32), "idx": idx, "box": np.zeros((n_struct, 3, 3)),
              "offsets": np.zeros((n_struct, p_max, 3)), "n_atoms": np.full((n_struct,), n, dtype=np.int32)}
    labels = {"energy": rng.normal(size=(n_struct,)), "forces": rng.normal(size=(n_struct, n, 3))}
    return inputs, labels
