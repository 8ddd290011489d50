"""NVT Nose-Hoover MD on the GPU against the NumPy restatement of jax_md's integrator (oracle/md_oracle.py)."""
import numpy as np
import pytest
import torch

from apax_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

MASS = {1: 1.008, 8: 15.999}


def test_nvt_steps_match_oracle():
    from apax_b200 import runtime as rt
    from apax_b200.utils.jax_md_reduced.simulate import nvt_nose_hoover
    from oracle import gmnn_oracle as go
    from oracle import md_oracle as mo
    from oracle import neighbor_oracle as no

    pos, Z, cell = syn.water_box(64, 3)
    box = cell.T.copy()
    frac = pos @ np.linalg.inv(cell)
    frac -= np.floor(frac)
    params = syn.gmnn_params(seed=4)
    mass = np.array([MASS[int(z)] for z in Z])
    rng = np.random.default_rng(0)
    kT, dt = 0.0259, 0.05                               # ~300 K in eV; dt in A sqrt(amu/eV) (~0.5 fs)
    p0 = rng.normal(size=(len(Z), 3)) * np.sqrt(mass * kT)[:, None]
    p0 -= p0.mean(0)
    pairs = no.jaxmd_sparse_pairs(frac, box, 6.0, 0.5)   # list with skin, valid for the few steps below

    def oracle_force(R):
        return go.energy_forces(params, R, Z, pairs, box, None, "minimage")["forces"]

    st = mo.nvt_init(frac, p0, mass, oracle_force, kT, dt)
    dp = rt.DeviceParams(params)
    prepared = rt.PreparedSystem(dp, Z, pairs)
    box_d = rt.to_device(box, torch.float64)

    def gpu_force(R, box=None, **kw):
        return prepared.compute(R, box_d, None, "minimage")[1][0]

    init_fn, apply_fn = nvt_nose_hoover(gpu_force, None, dt, kT)
    gst = init_fn(None, frac, mass=mass, momentum=p0)
    assert abs(gst.chain_numbers()["KE"] - st.chain.KE) <= 1e-12 * st.chain.KE
    for step in range(5):
        st = mo.nvt_step(st, oracle_force, box, kT, dt)
        gst = apply_fn(gst, box=box_d)
    torch.cuda.synchronize()
    ch = gst.chain_numbers()
    # fp32 force noise (~1e-6 relative) enters momenta through dt/2 F: compare at that level
    assert np.abs(gst.momentum.cpu().numpy() - st.momentum).max() <= 1e-5 * np.abs(st.momentum).max()
    d = gst.position.cpu().numpy() - st.position
    d -= np.round(d)
    assert np.abs(d).max() < 1e-8
    assert np.allclose(ch["xi"][:5], st.chain.xi, rtol=1e-6, atol=1e-12)
    assert np.allclose(ch["p_xi"][:5], st.chain.p_xi, rtol=1e-5, atol=1e-9)
    assert np.allclose(ch["Q"][:5], st.chain.Q, rtol=1e-7)
    assert abs(ch["KE"] - st.chain.KE) <= 1e-6 * st.chain.KE


def test_run_nvt_conserves_extended_energy_and_rebuilds():
    from apax_b200.md.ase_calc import Atoms
    from apax_b200.md.simulate import run_nvt
    from oracle import md_oracle as mo

    pos, Z, cell = syn.water_box(64, 5)
    params = syn.gmnn_params(seed=4)
    params["params"]["energy_model"]["scale_shift"]["scale_per_element"][:] = 0.02  # soften the random-init surface
    mass = np.array([MASS[int(z)] for z in Z])
    kT, dt = 0.0259, 0.02
    log = []

    def on_step(i, state):
        if i % 20 == 0:
            log.append(float(state.chain_numbers()["KE"]))

    state, counters = run_nvt(params, Atoms(pos, Z, cell), dt=dt, kT=kT, n_steps=100, masses=mass, rebuild_every=10,
                              on_step=on_step)
    assert counters["steps"] == 100 and counters["rebuilds"] == 10   # steps 10, 20, ..., 100
    p = state.momentum.cpu().numpy()
    assert np.isfinite(p).all() and np.isfinite(state.position.cpu().numpy()).all()
    assert np.abs(p.sum(0)).max() < 1e-6 * np.abs(p).sum()           # total momentum stays zero
    assert 0.0 <= state.position.min().item() and state.position.max().item() < 1.0   # wrapped fractional coordinates
    assert all(k > 0 for k in log)
    state2, counters2 = run_nvt(params, Atoms(pos, Z, cell), dt=dt, kT=kT, n_steps=30, masses=mass, dr_threshold=0.5)
    assert counters2["rebuilds"] >= 0 and np.isfinite(state2.momentum.cpu().numpy()).all()


def test_nvt_kernels_against_reference_fixture(golden):
    """The device integrator alone (analytic tether force computed with torch on the GPU) against the fixture
    produced by the reference's own simulate.py: positions/momenta/chain after 1 and 7 steps."""
    from apax_b200 import runtime as rt
    from apax_b200.utils.jax_md_reduced.simulate import nvt_nose_hoover

    d = golden("md_nvt")
    box = rt.to_device(d["cell"].T.copy(), torch.float64)
    anchor = rt.to_device(d["frac0"], torch.float64)
    k = float(d["kspring"])

    def force(R, box=None, **kw):
        dd = R - anchor
        dd = dd - torch.round(dd)
        return -k * (dd @ rt.to_device(d["cell"].T.copy(), torch.float64).T)

    init_fn, apply_fn = nvt_nose_hoover(force, None, float(d["dt"]), float(d["kT"]))
    st = init_fn(None, d["frac0"], mass=d["mass"], momentum=d["p0"])
    assert abs(st.chain_numbers()["KE"] - float(d["KE0"])) < 1e-12
    for step in range(1, 8):
        st = apply_fn(st, box=box)
        if step in (1, 7):
            ch = st.chain_numbers()
            assert np.abs(st.position.cpu().numpy() - d[f"position_{step}"]).max() < 1e-13
            assert np.abs(st.momentum.cpu().numpy() - d[f"momentum_{step}"]).max() < 1e-12
            assert np.abs(ch["xi"][:5] - d[f"xi_{step}"]).max() < 1e-13
            assert np.abs(ch["p_xi"][:5] - d[f"p_xi_{step}"]).max() < 1e-11
            assert np.array_equal(ch["Q"][:5], d[f"Q_{step}"])
            assert abs(ch["KE"] - float(d[f"KE_{step}"])) < 1e-11


def test_multi_step_call_follows_the_per_step_loop():
    """apax_md_nvt_run (all steps between two list builds in one library call) against the per-step host loop: same
    integrator and forces; the lists are rebuilt from positions one half-step apart, both valid thanks to the skin, so
    the trajectories agree to fp32 summation noise."""
    from apax_b200.md.ase_calc import Atoms
    from apax_b200.md.simulate import run_nvt

    pos, Z, cell = syn.water_box(64, 5)
    params = syn.gmnn_params(seed=4)
    params["params"]["energy_model"]["scale_shift"]["scale_per_element"][:] = 0.02
    mass = np.array([MASS[int(z)] for z in Z])
    rng = np.random.default_rng(3)
    p0 = rng.normal(size=(len(Z), 3)) * np.sqrt(mass * 0.0259)[:, None]
    p0 -= p0.mean(0)
    kw = dict(dt=0.02, kT=0.0259, n_steps=25, masses=mass, momentum=p0, rebuild_every=10)
    fast, c_fast = run_nvt(params, Atoms(pos, Z, cell), **kw)
    slow, c_slow = run_nvt(params, Atoms(pos, Z, cell), on_step=lambda i, s: None, **kw)
    assert c_fast["steps"] == c_slow["steps"] == 25 and c_fast["rebuilds"] == c_slow["rebuilds"] == 2
    dR = (fast.position - slow.position).cpu().numpy()
    dR -= np.round(dR)
    assert np.abs(dR).max() < 1e-7
    assert np.abs((fast.momentum - slow.momentum).cpu().numpy()).max() < 1e-5 * np.abs(slow.momentum.cpu().numpy()).max()
