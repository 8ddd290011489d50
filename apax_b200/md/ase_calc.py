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
        reference's padding to the largest structure).  Returns a list of result dicts.

        The batches are software-pipelined: batch b is enqueued completely (lists, E+F, ensemble statistics, copies into
        pinned staging buffers) before the host waits for batch b-1's copies and slices its results, so the host-side work
        of one batch hides behind the GPU work of the next.  No synchronisation happens inside a batch: the list is
        evaluated at its capacity (padding entries are (0,0) pairs, which contribute exactly nothing) and the overflow flag
        travels back with the results; an overflowing batch is simply redone with a larger capacity."""
        import torch

        model = self.builder.build_energy_derivative_model(init_box=self._cell(atoms_list[0]).T)
        em = getattr(model, "energy_model", model)
        dp = device_params(self.params, em.statics())
        r_max = float(dp.config["r_max"])
        n_ens = dp.n_out
        results = [None] * len(atoms_list)
        state = self.__dict__.setdefault("_batch_state", {"pairs_per_atom": None})

        def enqueue(start, chunk, cap, parity):
            pos = np.concatenate([np.asarray(a.positions, dtype=np.float64) for a in chunk])
            num = np.concatenate([np.asarray(a.numbers, dtype=np.int32) for a in chunk])
            cells = np.stack([self._cell(a) for a in chunk])
            ptr = np.concatenate([[0], np.cumsum([len(a.numbers) for a in chunk])]).astype(np.int32)
            idx, _, offsets, n_found, overflow = rt.nl_build_images_batched(pos, cells, ptr, r_max, cap)
            e, f = rt.energy_forces_batched(dp, pos, num, idx, offsets, ptr)       # whole capacity: the tail is (0,0) padding
            tag = f"p{parity}"
            if n_ens == 1:
                host = [self._to_host(tag + nm, t) for nm, t in zip(("nf", "ov", "e", "f"), (n_found, overflow, e, f))]
            else:
                stats = rt.ensemble_stats(e, f)
                host = [self._to_host(tag + nm, t) for nm, t in
                        zip(("nf", "ov", "e", "e_mean", "e_unc", "f_mean", "f_unc", "f_ens"), (n_found, overflow, e) + stats)]
            ev = torch.cuda.Event()
            ev.record()
            return {"start": start, "chunk": chunk, "ptr": ptr, "host": host, "event": ev, "cap": cap, "n_atoms": len(num)}

        def finalize(rec):
            rec["event"].synchronize()
            host = rec["host"]
            nf, ov = int(host[0].numpy()[0]), int(host[1].numpy()[0])
            if ov or nf > rec["cap"]:
                return int(nf * self.padding_factor) + 1024          # "neighbor list overflowed, extending."
            state["pairs_per_atom"] = max(state["pairs_per_atom"] or 0.0, nf / rec["n_atoms"])
            ptr, start = rec["ptr"], rec["start"]
            if n_ens == 1:
                e, f = host[2].numpy().copy(), host[3].numpy().copy()
                for b in range(len(rec["chunk"])):
                    results[start + b] = {"energy": float(e[b, 0]), "forces": f[0, ptr[b]:ptr[b + 1]]}
                return None
            e_h, e_mean, e_unc, fm, f_unc, f_ens = [h.numpy().copy() for h in host[2:]]
            for b in range(len(rec["chunk"])):
                sl = slice(ptr[b], ptr[b + 1])
                results[start + b] = {"energy": float(e_mean[b]), "energy_ensemble": e_h[b], "energy_uncertainty": float(e_unc[b]),
                                      "forces": fm[sl], "forces_uncertainty": f_unc[sl], "forces_ensemble": f_ens[sl]}
            return None

        def capacity(chunk):
            n = sum(len(a.numbers) for a in chunk)
            ppa = state["pairs_per_atom"]
            return int(n * self.max_pairs_per_atom) if ppa is None else int(n * ppa * 1.15) + 1024

        def run_sync(start, chunk, cap):
            while True:
                need = finalize(enqueue(start, chunk, cap, 0))
                if need is None:
                    return
                cap = need

        pending, parity = None, 0
        for s in range(0, len(atoms_list), batch_size):
            chunk = atoms_list[s:s + batch_size]
            if state["pairs_per_atom"] is None:          # first batch ever: size the later ones from its occupancy
                run_sync(s, chunk, capacity(chunk))
                continue
            rec = enqueue(s, chunk, capacity(chunk), parity)
            parity ^= 1
            if pending is not None:
                need = finalize(pending)
                if need is not None:
                    rec["event"].synchronize()           # drain, then redo the overflowing batch on its own
                    run_sync(pending["start"], pending["chunk"], need)
            pending = rec
        if pending is not None:
            need = finalize(pending)
            if need is not None:
                run_sync(pending["start"], pending["chunk"], need)
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
