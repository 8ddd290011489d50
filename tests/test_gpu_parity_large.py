"""Parity of the CUDA path against the oracle AT THE SIZES BASELINE.json NAMES: the 12,288-atom water box of configs[2]
(whole box, both the resident-list path and the host-buffer call apax_ctx_calculate) and the 1,000,002-atom box of
configs[3] (oracle-evaluated sub-blocks: the model is strictly one-hop, so the forces on the atoms inside a sphere of radius
r_in and their atomic energies depend only on the atoms within r_in + 2 r_max of its centre).

Energies: 1e-6 relative (north_star).  Forces: tests/conftest.py::assert_force_parity -- the margins against the
reference-precision oracle AND against the all-fp64 oracle are printed and stored (gpurun_out/force_parity_report.json)."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn
from conftest import assert_force_parity, record_parity

pytestmark = pytest.mark.gpu

R_MAX = 6.0


@pytest.fixture(scope="module")
def rt():
    from apax_b200 import runtime

    assert torch.cuda.is_available()
    return runtime


@pytest.fixture(scope="module")
def box_12288():
    from oracle import gmnn_oracle as go
    from oracle import neighbor_oracle as no

    pos, Z, cell = syn.water_12288()
    frac = pos @ np.linalg.inv(cell)
    idx = no.jaxmd_sparse_pairs_fast(frac, cell.T, R_MAX)          # the reference's pair set and order (partition.py:640-664)
    params = syn.gmnn_params(seed=1234)
    ref32 = go.energy_forces(params, frac, Z, idx, cell.T, None, "minimage")
    ref64 = go.energy_forces(params, frac, Z, idx, cell.T, None, "minimage", go.fp64_config())
    return dict(pos=pos, Z=Z, cell=cell, frac=frac, idx=idx, params=params, ref32=ref32, ref64=ref64)


def test_12288_resident_list_vs_oracle(rt, box_12288):
    """BASELINE configs[2]'s box, list built on the GPU (bit-exact pair set), E+F through apax_gm_prepare / apax_gm_compute."""
    d = box_12288
    n = len(d["Z"])
    idx, n_found, ov = rt.nl_build_minimage(d["frac"], d["cell"].T, R_MAX, int(n * 100))
    torch.cuda.synchronize()
    nf = int(n_found.item())
    assert not int(ov.item())
    assert np.array_equal(idx[:, :nf].cpu().numpy(), d["idx"])            # neighbour indices bit-exact, reference order
    dp = rt.DeviceParams(d["params"])
    system = rt.PreparedSystem(dp, d["Z"], idx[:, :nf].contiguous())
    e, f = system.compute(rt.to_device(d["frac"], torch.float64), rt.to_device(d["cell"].T, torch.float64), None, "minimage")
    torch.cuda.synchronize()
    e, f = e.cpu().numpy(), f.cpu().numpy()[0]
    rel = abs(e[0] - d["ref32"]["energy"]) / abs(d["ref32"]["energy"])
    record_parity("12288 resident list: energy", rel_err_vs_ref32=rel, rel_err_vs_fp64=abs(e[0] - d["ref64"]["energy"]) / abs(d["ref64"]["energy"]),
                  ref32_rel_err_vs_fp64=abs(d["ref32"]["energy"] - d["ref64"]["energy"]) / abs(d["ref64"]["energy"]))
    assert rel <= 1e-6, rel
    assert_force_parity(f, d["ref32"]["forces"], d["ref64"]["forces"], where="12288 resident list")
    # per-atom energies (models.py:109-122) add up to the total in the reference's fp64 sum
    ea = system.atomic_energies().cpu().numpy()
    assert ea.shape == (1, n) and abs(ea.sum() - e[0]) <= 1e-9 * abs(e[0])


def test_12288_host_context_vs_oracle(rt, box_12288):
    """The same box through the host-buffer call (apax_ctx_calculate: H2D, list build, E+F, D2H), single point and with the
    MD skin (list cutoff 6.5: the extra entries are beyond r_max and contribute exactly zero)."""
    d = box_12288
    n = len(d["Z"])
    dp = rt.DeviceParams(d["params"])
    ctx = rt.HostContext(dp, n, 130 * n)
    for skin in (0.0, 0.5):
        e, f = ctx.calculate(d["pos"], d["Z"].astype(np.int32), d["cell"], dr_threshold=skin, rebuild=True)
        rel = abs(e[0] - d["ref32"]["energy"]) / abs(d["ref32"]["energy"])
        assert rel <= 1e-6, (skin, rel)
        assert_force_parity(f[0], d["ref32"]["forces"], d["ref64"]["forces"], where=f"12288 apax_ctx_calculate skin {skin}")
    ctx.close()


def _sub_block_oracle(params, pos, Z, cell, centre_atom, r_in, cfg=None):
    """Oracle on the cluster around `centre_atom`: atoms within r_in + 2 r_max (min-image, unwrapped around the centre) as a
    gas-phase cluster; centrals = atoms within r_in + r_max; returns (interior atom ids, sum of their E_i, their forces)."""
    from scipy.spatial import cKDTree

    from oracle import gmnn_oracle as go

    L = np.diag(cell)
    assert np.allclose(cell, np.diag(L))
    d = pos - pos[centre_atom]
    d -= L * np.round(d / L)
    r = np.linalg.norm(d, axis=1)
    sel = np.nonzero(r < r_in + 2 * R_MAX)[0]
    xyz = d[sel]
    central = r[sel] < r_in + R_MAX
    interior = r[sel] < r_in
    tree = cKDTree(xyz)
    pairs = tree.query_pairs(R_MAX * (1 + 1e-9), output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1]])
    j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    keep = central[i]                                  # rows of central atoms only: idx[1] = central, idx[0] = neighbour
    idx = np.stack([j[keep], i[keep]]).astype(np.int64)
    cfg = cfg or go.GMNNConfig()
    R, Zt, it, bt, off = go._as_inputs(xyz, Z[sel], idx, np.zeros((3, 3)), None)
    R = R.clone().requires_grad_(True)
    _, aux = go.energy_model(params, R, Zt, it, bt, off, "gas", cfg, return_aux=True)
    E_i = aux["E_i"][:, 0]
    e_central = E_i[torch.as_tensor(central)].sum()
    (grad,) = torch.autograd.grad(e_central, R)
    ids = sel[interior]
    return ids, float(E_i[torch.as_tensor(interior)].sum().detach()), (-grad).numpy()[interior]


def test_1m_atom_box_sub_blocks_vs_oracle(rt):
    """BASELINE configs[3]: one E+F evaluation of the 1,000,002-atom box on the GPU; eight spheres of radius 10 A (~420 atoms
    each) are re-evaluated by the oracle in reference precision and in fp64 with everything they can see (22 A, ~4,500 atoms):
    per-atom forces of the interior atoms and the sum of their atomic energies."""
    from oracle import gmnn_oracle as go

    pos, Z, cell = syn.water_1m()
    n = len(Z)
    assert n == 1_000_002
    params = syn.gmnn_params(seed=1234)
    dp = rt.DeviceParams(params)
    frac = rt.to_device(pos @ np.linalg.inv(cell), torch.float64)
    box = rt.to_device(cell.T.copy(), torch.float64)
    idx, n_found, ov = rt.nl_build_minimage(frac, box, R_MAX, int(n * 96) + 4096)
    torch.cuda.synchronize()
    nf = int(n_found.item())
    assert not int(ov.item())
    system = rt.PreparedSystem(dp, Z, idx[:, :nf].contiguous())
    del idx
    e, f = system.compute(frac, box, None, "minimage")
    ea = system.atomic_energies()
    torch.cuda.synchronize()
    f = f.cpu().numpy()[0]
    ea = ea.cpu().numpy()[0]
    assert np.isfinite(f).all() and np.isfinite(ea).all()
    assert abs(ea.sum() - float(e.cpu().numpy()[0])) <= 1e-9 * abs(ea.sum())
    assert np.abs(f.sum(0)).max() <= 1e-4 * np.abs(f).max() * np.sqrt(n)          # Newton's third law up to fp32 pair rounding
    rng = np.random.default_rng(42)
    centres = rng.choice(n, size=8, replace=False)
    f_k, f_32, f_64, e_rel = [], [], [], []
    for c in centres:
        ids, e32, f32 = _sub_block_oracle(params, pos, Z, cell, int(c), 10.0)
        ids64, e64, f64 = _sub_block_oracle(params, pos, Z, cell, int(c), 10.0, go.fp64_config())
        assert np.array_equal(ids, ids64) and len(ids) > 300
        ek = ea[ids].sum()
        e_rel.append(abs(ek - e32) / abs(e32))
        assert abs(ek - e32) <= 1e-6 * abs(e32), (c, ek, e32)
        f_k.append(f[ids]); f_32.append(f32); f_64.append(f64)
    record_parity("1M box sub-blocks: energy", max_rel_err_vs_ref32=max(e_rel), blocks=len(centres))
    assert_force_parity(np.concatenate(f_k), np.concatenate(f_32), np.concatenate(f_64), where="1M box sub-blocks (8 x r_in 10 A)")
    del system
    torch.cuda.empty_cache()
