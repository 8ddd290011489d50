"""Host mirror of apax/md/ase_calc.py::ASECalculator for the energy+forces properties.  ASE itself
is not a dependency: `calculate` accepts any object with `.positions`, `.numbers`, `.cell` (3x3 array or
an object with `.array`) -- an `ase.Atoms` qualifies.  The per-call flow (neighbour path selection
:501-522, list update, model call, results to NumPy fp64 :392-393) is one apax_ctx_calculate call."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .. import runtime as rt
from ..nn.models import device_params


@dataclass
class Atoms:
    """Minimal stand-in with ase.Atoms' attribute names."""
    positions: np.ndarray
    numbers: np.ndarray
    cell: np.ndarray
    pbc: bool = True


class ASECalculator:
    implemented_properties = ["energy", "forces"]   # + "stress" when the model config sets calc_stress (ase_calc.py:180-181)

    def __init__(self, params: dict, model_config: Optional[dict] = None, dr_threshold: float = 0.5,
                 padding_factor: float = 1.5, max_pairs_per_atom: int = 160):
        from ..nn.builder import GMNNBuilder

        self.params = params
        self.builder = GMNNBuilder(model_config)
        self.dr_threshold = dr_threshold
        self.padding_factor = padding_factor
        self.max_pairs_per_atom = max_pairs_per_atom
        self.ctx: Optional[rt.HostContext] = None
        self.results: dict = {}
        self.numbers: Optional[np.ndarray] = None
        self.implemented_properties = ["energy", "forces"] + (["stress"] if self.builder.config.get("calc_stress") else [])

    def _cell(self, atoms) -> np.ndarray:
        cell = atoms.cell
        return np.asarray(getattr(cell, "array", cell), dtype=np.float64)

    def initialize(self, atoms, max_pairs: Optional[int] = None):
        model = self.builder.build_energy_derivative_model(init_box=self._cell(atoms).T)
        em = getattr(model, "energy_model", model)
        self.dp = device_params(self.params, em.statics())
        n = len(atoms.numbers)
        self.max_pairs = max_pairs or n * self.max_pairs_per_atom
        if self.ctx is not None:
            self.ctx.close()
        self.ctx = rt.HostContext(self.dp, n, self.max_pairs)
        self.n_atoms = n
        self.numbers = np.array(atoms.numbers, dtype=np.int32)

    def calculate(self, atoms, properties=("energy",), system_changes=None):
        # re-initialise on ANY change of the atomic numbers, not only of their count (ase_calc.py:366-370)
        if self.ctx is None or len(atoms.numbers) != self.n_atoms or not np.array_equal(self.numbers, atoms.numbers):
            self.initialize(atoms)
        cell = self._cell(atoms)
        try:
            e, f = self.ctx.calculate(atoms.positions, atoms.numbers, cell, dr_threshold=self.dr_threshold)
        except rt._lib.ApaxB200Error as err:
            if err.code != rt._lib.ERR_OVERFLOW:
                raise
            # "neighbor list overflowed, reallocating." (ase_calc.py:381-384, 345-347)
            self.initialize(atoms, max_pairs=int(self.ctx.n_pairs * self.padding_factor) + 1024)
            e, f = self.ctx.calculate(atoms.positions, atoms.numbers, cell, dr_threshold=self.dr_threshold)
        if self.dp.n_out == 1:
            self.results = {"energy": float(e[0]), "forces": f[0]}
            if "stress" in self.implemented_properties:      # ProcessStress (function_transformations.py:140-159)
                self.results["stress"] = self.ctx.stress()
        else:
            n_ens = self.dp.n_out
            fm = f.mean(axis=0)
            self.results = {"energy": float(e.mean()), "energy_ensemble": e, "energy_uncertainty": float(np.sqrt(((e - e.mean()) ** 2).sum() / (n_ens - 1))),
                            "forces": fm, "forces_uncertainty": np.sqrt(((f - fm) ** 2).sum(axis=0) / (n_ens - 1)),
                            "forces_ensemble": np.transpose(f, (1, 2, 0))}
        return self.results

    def batch_eval(self, atoms_list, batch_size: int = 64, silent: bool = True):
        """ASECalculator.batch_eval (apax/md/ase_calc.py:395-481): evaluate many structures, `batch_size` at a
        time, each batch in ONE neighbour-list call and ONE energy+forces call (atoms concatenated instead of the
        reference's padding to the largest structure).  Returns a list of result dicts."""
        import torch

        model = self.builder.build_energy_derivative_model(init_box=self._cell(atoms_list[0]).T)
        em = getattr(model, "energy_model", model)
        dp = device_params(self.params, em.statics())
        r_max = float(dp.config["r_max"])
        results = []
        for s in range(0, len(atoms_list), batch_size):
            chunk = atoms_list[s:s + batch_size]
            pos = np.concatenate([np.asarray(a.positions, dtype=np.float64) for a in chunk])
            num = np.concatenate([np.asarray(a.numbers, dtype=np.int32) for a in chunk])
            cells = np.stack([self._cell(a) for a in chunk])
            ptr = np.concatenate([[0], np.cumsum([len(a.numbers) for a in chunk])]).astype(np.int32)
            cap = int(len(num) * self.max_pairs_per_atom)
            while True:
                idx, _, offsets, n_found, overflow = rt.nl_build_images_batched(pos, cells, ptr, r_max, cap)
                nf = int(n_found.item())
                if not int(overflow.item()):
                    break
                cap = int(nf * self.padding_factor) + 1024      # "neighbor list overflowed, extending."
            e, f = rt.energy_forces_batched(dp, pos, num, idx[:, :nf].contiguous(), offsets[:nf].contiguous(), ptr)
            n_ens = dp.n_out
            if n_ens == 1:
                torch.cuda.synchronize()
                e, f = e.cpu().numpy(), f.cpu().numpy()
                for b in range(len(chunk)):
                    results.append({"energy": float(e[b, 0]), "forces": f[0, ptr[b]:ptr[b + 1]]})
                continue
            # ensemble statistics (models.py:221-263) for the whole chunk in one kernel pair on the device, results through
            # pinned staging buffers, sliced per structure afterwards
            stats = rt.ensemble_stats(e, f)
            host = [self._to_host(nm, t) for nm, t in zip(("e", "e_mean", "e_unc", "f_mean", "f_unc", "f_ens"), (e,) + stats)]
            torch.cuda.synchronize()
            e_h, e_mean, e_unc, fm, f_unc, f_ens = [h.numpy().copy() for h in host]
            for b in range(len(chunk)):
                sl = slice(ptr[b], ptr[b + 1])
                results.append({"energy": float(e_mean[b]), "energy_ensemble": e_h[b], "energy_uncertainty": float(e_unc[b]),
                                "forces": fm[sl], "forces_uncertainty": f_unc[sl], "forces_ensemble": f_ens[sl]})
        return results

    def _to_host(self, name: str, t):
        """Device tensor -> pinned host staging buffer (cached per name and shape), asynchronously on the current stream."""
        import torch

        cache = self.__dict__.setdefault("_pinned", {})
        key = (name, tuple(t.shape), t.dtype)
        if key not in cache:
            cache[key] = torch.empty(tuple(t.shape), dtype=t.dtype).pin_memory()
        cache[key].copy_(t, non_blocking=True)
        return cache[key]

    def get_potential_energy(self, atoms):
        return self.calculate(atoms)["energy"]

    def get_forces(self, atoms):
        return self.calculate(atoms)["forces"]
