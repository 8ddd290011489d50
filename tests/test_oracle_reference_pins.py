"""The reference's own relational tests for this path, replayed on the oracle."""
import numpy as np
import torch

from apax_b200 import synthetic as syn
from oracle import gmnn_oracle as go
from oracle import neighbor_oracle as no


def test_scale_shift_numbers():
    # tests/unit_tests/layers/test_scaling.py:7-41
    x = torch.ones(3, 1)
    Z = torch.tensor([1, 2, 2])
    cfg = go.GMNNConfig()
    out = go.per_element_scale_shift(x, Z, torch.full((119, 1), 1.0, dtype=torch.float64),
                                     torch.full((119, 1), 2.0, dtype=torch.float64), cfg)
    assert np.allclose(out.numpy(), [[3.0], [3.0], [3.0]])
    out = go.per_element_scale_shift(x, Z, torch.tensor([[10.0], [2.0], [3.0]], dtype=torch.float64),
                                     torch.tensor([[0.0], [-2.0], [2.0]], dtype=torch.float64), cfg)
    assert np.allclose(out.numpy(), [[0.0], [5.0], [5.0]])


def test_triangular_index_sets():
    # apax/layers/descriptor/triangular_indices.py:28-49; the 3-d set is NOT the sorted i<=j<=k set
    assert len(go.tril_2d_indices(5)) == 15
    t3 = go.tril_3d_indices(5)
    assert len(t3) == 35
    assert t3[:6].tolist() == [[0, 0, 0], [0, 1, 1], [0, 2, 2], [0, 3, 3], [0, 4, 4], [0, 1, 2]]
    assert [1, 0, 0] in t3.tolist()


def test_invariances_and_finite_differences():
    pos, Z, cell = syn.water_box(8, 2)
    p = syn.gmnn_params(seed=5)
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
    frac = pos @ np.linalg.inv(cell)
    cfg64 = go.fp64_config()
    r = go.energy_forces(p, frac, Z, idx, cell.T, off, "offsets", cfg64)
    assert np.abs(r["forces"].sum(0)).max() < 1e-9          # translation invariance
    # permutation of atoms
    perm = np.random.default_rng(0).permutation(len(Z))
    inv = np.argsort(perm)
    idx_p = inv[idx]
    rp = go.energy_forces(p, frac[perm], Z[perm], idx_p, cell.T, off, "offsets", cfg64)
    assert abs(rp["energy"] - r["energy"]) < 1e-9
    assert np.abs(rp["forces"] - r["forces"][perm]).max() < 1e-9
    # rotation of the whole system (positions and cell)
    a = 0.7
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    cell_r = cell @ rot.T
    idx_r, off_r = no.vesin_idx_offsets(pos @ rot.T, cell_r, 6.0)
    rr = go.energy_forces(p, frac, Z, idx_r, cell_r.T, off_r, "offsets", cfg64)
    assert abs(rr["energy"] - r["energy"]) < 1e-8
    # central finite differences on Cartesian displacements
    h = 1e-5
    icell = np.linalg.inv(cell)
    for a_i, k in [(0, 0), (3, 2), (10, 1)]:
        dc = np.zeros(3)
        dc[k] = h
        fp, fm = frac.copy(), frac.copy()
        fp[a_i] += dc @ icell
        fm[a_i] -= dc @ icell
        ep = go.energy_forces(p, fp, Z, idx, cell.T, off, "offsets", cfg64)["energy"]
        em = go.energy_forces(p, fm, Z, idx, cell.T, off, "offsets", cfg64)["energy"]
        assert abs(-(ep - em) / (2 * h) - r["forces"][a_i, k]) < 1e-6


def test_min_image_equals_multi_image_when_box_is_large():
    # cross-path equality in the spirit of tests/unit_tests/md/test_openmm_interface_against_ase.py
    pos, Z, cell = syn.water_box(64, 1)  # L = 12.4 > 2 r_max
    p = syn.gmnn_params(seed=3)
    frac = pos @ np.linalg.inv(cell)
    pairs = no.jaxmd_sparse_pairs(frac, cell.T, 6.0)
    r1 = go.energy_forces(p, frac, Z, pairs, cell.T, None, "minimage")
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
    assert idx.shape[1] == pairs.shape[1]
    r2 = go.energy_forces(p, frac, Z, idx, cell.T, off, "offsets")
    assert abs(r1["energy"] - r2["energy"]) < 1e-6 * abs(r1["energy"])
    assert np.abs(r1["forces"] - r2["forces"]).max() < 2e-6


def test_vesin_restatement_distance_multiset():
    # tests/unit_tests/data/test_input_pipeline.py:150-212: small cubic / orthorhombic / triclinic cells
    cells = [np.array([[1.8, 0.1, 0.0], [0.0, 2.5, 0.1], [0.1, 0.0, 2.5]]),
             np.array([[1.8, 0.0, 0.0], [0.0, 2.5, 0.0], [0.0, 0.0, 2.5]]),
             np.array([[1.5, 0.0, 0.5], [0.0, 2.5, 0.0], [0.0, 0.5, 1.5]])]
    pos = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
    for cell in cells:
        ii, jj, ss = no.vesin_full_list(pos, cell, 2.0)
        d = np.linalg.norm(pos[jj] - pos[ii] + ss @ cell, axis=1)
        assert np.all(d < 2.0) and len(d) > 0
        # brute-force count over a generous image range
        cnt = 0
        for s in np.ndindex(9, 9, 9):
            S = np.array(s) - 4
            for i in range(2):
                for j in range(2):
                    if i == j and not S.any():
                        continue
                    cnt += np.linalg.norm(pos[j] - pos[i] + S @ cell) < 2.0
        assert cnt == len(d)
        # symmetric: (i,j,S) <-> (j,i,-S)
        fwd = set(zip(ii.tolist(), jj.tolist(), map(tuple, ss.tolist())))
        assert all((j, i, tuple(-np.array(s))) in fwd for i, j, s in fwd)


def test_fast_pair_search_equals_brute_force():
    pos, Z, cell = syn.water_box(400, 11)
    frac = pos @ np.linalg.inv(cell) + np.array([0.3, -1.2, 2.0])  # also unwrapped coordinates
    a = no.jaxmd_sparse_pairs(frac, cell.T, 6.0, 0.5)
    b = no.jaxmd_sparse_pairs_fast(frac, cell.T, 6.0, 0.5)
    assert np.array_equal(a, b)
