"""Neighbour lists built on the GPU against the oracle: pair sets bit-exact, reference order."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt():
    from apax_b200 import runtime

    return runtime


@pytest.fixture(scope="module")
def no():
    from oracle import neighbor_oracle

    return neighbor_oracle


def test_min_image_list_matches_reference_fixture(rt, golden):
    d = golden("water_min_image")
    cap = d["idx"].shape[1]
    idx, n_found, ov = rt.nl_build_minimage(d["frac"], d["box"], 6.0 + float(d["skin"]), cap)
    assert np.array_equal(idx.cpu().numpy(), d["idx"])       # indices, order and N-padding, bit for bit
    assert int(n_found.item()) == int(d["n_valid"]) and int(ov.item()) == 0
    idx2, _, _ = rt.nl_build_minimage(d["moved_big"], d["box"], 6.0 + float(d["skin"]), cap)
    assert np.array_equal(idx2.cpu().numpy(), d["idx_big"])
    # skin test (partition.py:771-779)
    assert int(rt.nl_needs_rebuild(d["moved_small"], d["frac"], d["box"], float(d["skin"])).item()) == int(d["rebuilt_small"])
    assert int(rt.nl_needs_rebuild(d["moved_big"], d["frac"], d["box"], float(d["skin"])).item()) == int(d["rebuilt_big"])


def test_min_image_overflow_flag(rt, golden):
    d = golden("water_min_image")
    cap = int(d["n_valid"]) - 100
    idx, n_found, ov = rt.nl_build_minimage(d["frac"], d["box"], 6.0 + float(d["skin"]), cap)
    assert int(ov.item()) == 1 and int(n_found.item()) == int(d["n_valid"])
    assert np.array_equal(idx.cpu().numpy(), d["idx"][:, :cap])  # truncated exactly like idx[:, :max_occupancy]


@pytest.mark.parametrize("n_mol,seed,cutoff", [(64, 2, 6.0), (512, 3, 6.5), (1000, 4, 5.0)])
def test_min_image_list_vs_oracle(rt, no, n_mol, seed, cutoff):
    pos, Z, cell = syn.water_box(n_mol, seed)
    frac = pos @ np.linalg.inv(cell)
    ref = no.jaxmd_sparse_pairs(frac, cell.T, cutoff)
    cap = int(ref.shape[1] * 1.25)
    idx, n_found, ov = rt.nl_build_minimage(frac, cell.T, cutoff, cap)
    nf = int(n_found.item())
    assert nf == ref.shape[1] and not int(ov.item())
    assert np.array_equal(idx[:, :nf].cpu().numpy(), ref)
    assert np.all(idx[:, nf:].cpu().numpy() == len(Z))


def test_min_image_rows_longer_than_the_scratch_stride(rt, no):
    """Inhomogeneous density: a compact cluster inside a dilute box.  The single-scan build parks at most
    2*capacity/N neighbours per atom; the cluster's rows are several times longer and take the re-scan path, the
    dilute atoms stay on the fast path -- the list must still equal the oracle's, order included."""
    rng = np.random.default_rng(11)
    L = 40.0
    cell = np.diag([L, L, L])
    cluster = 20.0 + rng.uniform(-3.0, 3.0, size=(300, 3))        # ~300 mutual neighbours within 6.5 A
    dilute = rng.uniform(0.0, L, size=(2700, 3))
    pos = np.concatenate([cluster, dilute])[rng.permutation(3000)]
    frac = pos @ np.linalg.inv(cell)
    ref = no.jaxmd_sparse_pairs(frac, cell.T, 6.5)
    rows = np.bincount(ref[0], minlength=3000)
    cap = int(ref.shape[1] * 1.1)
    assert rows.max() > 2 * cap // 3000 + 32 and np.median(rows) < cap // 3000   # both paths are exercised
    idx, n_found, ov = rt.nl_build_minimage(frac, cell.T, 6.5, cap)
    nf = int(n_found.item())
    assert nf == ref.shape[1] and not int(ov.item())
    assert np.array_equal(idx[:, :nf].cpu().numpy(), ref)
    # a too-small capacity raises the overflow flag and reports the size needed (partition.py:759-768: the caller
    # must re-allocate); every row that fits entirely is still the reference's
    cap2 = ref.shape[1] - 1234
    idx2, n2, ov2 = rt.nl_build_minimage(frac, cell.T, 6.5, cap2)
    assert int(ov2.item()) == 1 and int(n2.item()) == ref.shape[1]
    whole = int(np.searchsorted(ref[1], ref[1, cap2 - 1]))   # first slot of the row cut by the capacity
    assert np.array_equal(idx2.cpu().numpy()[:, :whole], ref[:, :whole])


def test_min_image_triclinic_and_unwrapped_positions(rt, no):
    rng = np.random.default_rng(5)
    cell = np.array([[26.0, 0.0, 0.0], [3.0, 25.0, 0.0], [-2.0, 4.0, 27.0]])
    n = 1500
    frac = rng.uniform(-1.5, 2.5, size=(n, 3))  # deliberately outside [0,1)
    box = cell.T
    ref = no.jaxmd_sparse_pairs(frac, box, 6.0)
    idx, n_found, ov = rt.nl_build_minimage(frac, box, 6.0, int(ref.shape[1] * 1.3))
    nf = int(n_found.item())
    assert nf == ref.shape[1]
    assert np.array_equal(idx[:, :nf].cpu().numpy(), ref)


def _triples(idx_i, idx_j, S):
    return set(zip(idx_i.tolist(), idx_j.tolist(), map(tuple, np.asarray(S).tolist())))


@pytest.mark.parametrize("case", ["water96", "cell128", "tiny_triclinic", "large", "open"])
def test_image_list_vs_oracle(rt, no, case):
    pbc = True
    if case == "water96":
        pos, Z, cell = syn.water_96()
        cutoff = 6.0
    elif case == "cell128":
        pos, Z, cell = syn.cell_128(17)
        cutoff = 6.0
    elif case == "tiny_triclinic":
        cell = np.array([[1.5, 0.0, 0.5], [0.0, 2.5, 0.0], [0.0, 0.5, 1.5]])  # test_input_pipeline.py:165
        pos = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
        cutoff = 2.0
    elif case == "large":
        pos, Z, cell = syn.water_box(300, 8)
        pos = pos + 40.0 * np.random.default_rng(0).integers(-2, 3, size=pos.shape)  # atoms far outside the cell
        cutoff = 6.0
    else:
        R, _ = syn.ethanol_batch(1, 0)
        pos, cell, pbc, cutoff = R[0], np.eye(3) * 20.0, False, 6.0
    ii, jj, ss = no.vesin_full_list(pos, cell, cutoff, periodic=pbc)
    cap = int(len(ii) * 1.5) + 8
    idx, shifts, offsets, n_found, ov = rt.nl_build_images(pos, cell, cutoff, cap, pbc=pbc)
    nf = int(n_found.item())
    assert nf == len(ii) and not int(ov.item())
    idx_h, sh_h, off_h = idx.cpu().numpy(), shifts.cpu().numpy(), offsets.cpu().numpy()
    assert _triples(idx_h[0, :nf], idx_h[1, :nf], sh_h[:nf]) == _triples(ii, jj, ss)  # pair SET bit-exact
    assert np.array_equal(off_h[:nf], sh_h[:nf].astype(np.float64) @ cell)
    assert np.all(idx_h[:, nf:] == 0) and np.all(off_h[nf:] == 0)                     # (0,0) padding
    d = np.linalg.norm(pos[idx_h[1, :nf]] - pos[idx_h[0, :nf]] + off_h[:nf], axis=1)
    assert np.all(d < cutoff)


def test_host_context_both_paths(rt):
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no
    from conftest import assert_force_parity, force_tolerance_ok, max_force_violation

    p = syn.gmnn_params(seed=1234)
    dp = rt.DeviceParams(p)
    # images path (96 atoms, L < 2 r_max) = BASELINE config 1
    pos, Z, cell = syn.water_96()
    ctx = rt.HostContext(dp, 4096, 400000)
    e, f = ctx.calculate(pos, Z, cell)
    idx, off = no.vesin_idx_offsets(pos, cell, 6.0)
    assert ctx.n_pairs == idx.shape[1]
    ref = go.energy_forces(p, pos @ np.linalg.inv(cell), Z, idx, cell.T, off, "offsets")
    ref64_a = go.energy_forces(p, pos @ np.linalg.inv(cell), Z, idx, cell.T, off, "offsets", go.fp64_config())
    assert abs(e[0] - ref["energy"]) <= 1e-6 * abs(ref["energy"])
    assert_force_parity(f[0], ref["forces"], ref64_a["forces"], where="ctx image path, 96 atoms")
    # min-image path with skin (1536 atoms), twice: second call reuses the list
    pos, Z, cell = syn.water_box(512, 7)
    e, f = ctx.calculate(pos, Z, cell, dr_threshold=0.5)
    frac = pos @ np.linalg.inv(cell)
    pairs = no.jaxmd_sparse_pairs(frac, cell.T, 6.0, 0.5)
    assert ctx.n_pairs == pairs.shape[1]
    ref = go.energy_forces(p, frac, Z, pairs, cell.T, None, "minimage")
    ref64 = go.energy_forces(p, frac, Z, pairs, cell.T, None, "minimage", go.fp64_config())
    assert abs(e[0] - ref["energy"]) <= 1e-6 * abs(ref["energy"])
    assert_force_parity(f[0], ref["forces"], ref64["forces"], where="ctx min-image")
    pos2 = pos + np.random.default_rng(0).normal(scale=0.01, size=pos.shape)
    e2, f2 = ctx.calculate(pos2, Z, cell, dr_threshold=0.5)
    ref2 = go.energy_forces(p, pos2 @ np.linalg.inv(cell), Z, pairs, cell.T, None, "minimage")
    assert abs(e2[0] - ref2["energy"]) <= 1e-6 * abs(ref2["energy"])
    ref2_64 = go.energy_forces(p, pos2 @ np.linalg.inv(cell), Z, pairs, cell.T, None, "minimage", go.fp64_config())
    assert_force_parity(f2[0], ref2["forces"], ref2_64["forces"], where="ctx min-image, reused list")
    ctx.close()
