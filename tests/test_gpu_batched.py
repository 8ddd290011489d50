"""Batched evaluation of many small periodic cells (BASELINE configs[4]: shallow-ensemble inference on
independent 128-atom cells): batched image lists and one batched E+F call against the oracle per cell."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn
from conftest import assert_energy_parity, assert_force_parity

pytestmark = pytest.mark.gpu


def _batch(n_cells, seed0=0):
    pos, Z, cells, ptr = [], [], [], [0]
    for s in range(n_cells):
        p, z, c = syn.cell_128(seed0 + s)
        if s % 3 == 1:  # ragged batch: drop a few atoms; also a sheared cell
            p, z = p[:-5], z[:-5]
            c = c + np.array([[0, 0.4, 0], [0, 0, 0.3], [0.2, 0, 0]])
        pos.append(p)
        Z.append(z)
        cells.append(c)
        ptr.append(ptr[-1] + len(z))
    return np.concatenate(pos), np.concatenate(Z), np.stack(cells), np.array(ptr, dtype=np.int32), pos, Z


def test_batched_image_lists_and_ensemble_forces():
    from apax_b200 import runtime as rt
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no

    n_cells, n_ens = 6, 4
    R, Zc, cells, ptr, pos_l, Z_l = _batch(n_cells)
    cap = int(len(Zc) * 110)
    idx, shifts, offsets, n_found, ov = rt.nl_build_images_batched(R, cells, ptr, 6.0, cap)
    nf = int(n_found.item())
    assert not int(ov.item())
    idx_h, sh_h, off_h = idx[:, :nf].cpu().numpy(), shifts[:nf].cpu().numpy(), offsets[:nf].cpu().numpy()
    total = 0
    for b in range(n_cells):
        ii, jj, ss = no.vesin_full_list(pos_l[b], cells[b], 6.0)
        sel = (idx_h[1] >= ptr[b]) & (idx_h[1] < ptr[b + 1])
        got = set(zip((idx_h[0, sel] - ptr[b]).tolist(), (idx_h[1, sel] - ptr[b]).tolist(), map(tuple, sh_h[sel].tolist())))
        assert got == set(zip(ii.tolist(), jj.tolist(), map(tuple, ss.tolist())))        # pair set bit-exact per cell
        assert np.array_equal(off_h[sel], sh_h[sel].astype(np.float64) @ cells[b])
        total += len(ii)
    assert total == nf and np.all(np.diff(idx_h[1]) >= 0)

    params = syn.gmnn_params(seed=21, n_out=n_ens, random_scale_shift=True, random_bias=True)
    dp = rt.DeviceParams(params)
    e, f = rt.energy_forces_batched(dp, R, Zc, idx[:, :nf].contiguous(), offsets[:nf].contiguous(), ptr)
    e, f = e.cpu().numpy(), f.cpu().numpy()
    assert e.shape == (n_cells, n_ens) and f.shape == (n_ens, len(Zc), 3)
    for b in range(n_cells):
        i_b, off_b = no.vesin_idx_offsets(pos_l[b], cells[b], 6.0)
        frac = pos_l[b] @ np.linalg.inv(cells[b])
        ref = go.shallow_ensemble(params, frac, Z_l[b], i_b, cells[b].T, off_b, "offsets")
        ref64 = go.shallow_ensemble(params, frac, Z_l[b], i_b, cells[b].T, off_b, "offsets", go.fp64_config())
        assert_energy_parity(e[b], ref["energy_ensemble"], ref64["energy_ensemble"], where=f"batched cell {b}")
        f_b = np.transpose(f[:, ptr[b]:ptr[b + 1]], (1, 2, 0))
        assert_force_parity(f_b, ref["forces_ensemble"], ref64["forces_ensemble"], where=f"cell {b}")


def test_batch_eval_mirror():
    from apax_b200.md.ase_calc import ASECalculator, Atoms

    params = syn.gmnn_params(seed=21, n_out=4, random_scale_shift=True, random_bias=True)
    calc = ASECalculator(params, model_config={"ensemble": {"kind": "shallow", "n_members": 4, "force_variance": True}})
    atoms = [Atoms(*syn.cell_128(s)) for s in range(5)]
    out = calc.batch_eval(atoms, batch_size=3)
    assert len(out) == 5
    single = ASECalculator(params, model_config={"ensemble": {"kind": "shallow", "n_members": 4, "force_variance": True}})
    for a, r in zip(atoms, out):
        s = single.calculate(a)
        # same model, different (equally valid) neighbour order inside a row => fp32 summation order differs
        assert abs(r["energy"] - s["energy"]) <= 1e-6 * abs(s["energy"])
        assert_force_parity(r["forces"], s["forces"], where="batch_eval vs calculate")      # literal tolerance
        assert r["forces_ensemble"].shape == (128, 3, 4) and np.abs(r["forces_uncertainty"] - s["forces_uncertainty"]).max() < 1e-5


def test_batched_image_lists_skewed_cells_and_cutoff_shell():
    """The batched scan decides most candidates in fp32 and only the shell around the cutoff in fp64: the pair set must
    stay bit-exact for strongly sheared small cells (image range 2-3 per axis), atoms far outside their cell (non-zero
    wrap), and pairs placed within 1e-9 A on either side of the cutoff."""
    from apax_b200 import runtime as rt
    from oracle import neighbor_oracle as no

    rng = np.random.default_rng(11)
    pos_l, cells, ptr = [], [], [0]
    for s in range(5):
        L = [3.1, 4.0, 5.5, 7.3, 12.5][s]
        c = np.diag(rng.uniform(0.9, 1.3, 3) * L) + rng.uniform(-0.35, 0.35, (3, 3)) * L * (s % 2 + 0.2)
        n = [5, 9, 14, 23, 40][s]
        p = rng.uniform(-1.5, 2.5, (n, 3)) @ c                     # fractional coordinates in [-1.5, 2.5)
        if s == 4:                                                 # partners straddling the cutoff by +-1e-9 A
            u = rng.normal(size=(6, 3))
            u /= np.linalg.norm(u, axis=1)[:, None]
            for k in range(6):
                p[2 * k + 1] = p[2 * k] + u[k] * (6.0 + (1e-9 if k % 2 else -1e-9)) + np.array([1, -1, 0]) @ c * (k % 3 - 1)
        pos_l.append(p)
        cells.append(c)
        ptr.append(ptr[-1] + n)
    R, cells_a, ptr_a = np.concatenate(pos_l), np.stack(cells), np.array(ptr, dtype=np.int32)
    cap = 4000 * len(R)
    idx, shifts, offsets, n_found, ov = rt.nl_build_images_batched(R, cells_a, ptr_a, 6.0, cap)
    nf = int(n_found.item())
    assert not int(ov.item())
    idx_h, sh_h = idx[:, :nf].cpu().numpy(), shifts[:nf].cpu().numpy()
    total = 0
    for b in range(5):
        ii, jj, ss = no.vesin_full_list(pos_l[b], cells[b], 6.0)
        sel = (idx_h[1] >= ptr[b]) & (idx_h[1] < ptr[b + 1])
        got = set(zip((idx_h[0, sel] - ptr[b]).tolist(), (idx_h[1, sel] - ptr[b]).tolist(), map(tuple, sh_h[sel].tolist())))
        want = set(zip(ii.tolist(), jj.tolist(), map(tuple, ss.tolist())))
        assert got == want, (b, len(got), len(want))
        total += len(ii)
    assert total == nf
