"""Multi-GPU paths on a box with at least two GPUs (skipped otherwise; the CPU twins are tests/test_domain_gloo.py and
tests/test_train_dp_gloo.py): the slab-decomposed E+F with the halo exchange over NVLink peer memory against a single-GPU
evaluation of the same box, and the data-parallel training step with the fused all-reduce + Adam kernel against NCCL."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs_two = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")


def _torchrun(port, *bench_args, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", *bench_args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@needs_two
@pytest.mark.parametrize("extra", [(), ("--nccl",)])
def test_slab_decomposition_matches_single_gpu(extra):
    line = _torchrun(29611 + len(extra), "--molecules", "20000", "--steps", "2", "--warmup", "3", "--no-others", *extra)
    assert line["n_gpus"] == 2
    # every N > 1 line carries the comparison with a single-GPU evaluation of the whole box (same kernels, other summation
    # order in the fp64 gathers only): energies to fp64 rounding, forces inside the literal tolerance
    assert line["verify"]["energy_rel_err"] <= 1e-8
    assert line["verify"]["max_force_violation_x_tolerance"] <= 1.0
    assert line["roofline"]["frac"] > 0 and line["e2e"]["value"] > 0


@needs_two
def test_default_multi_gpu_line_carries_the_other_workloads():
    """The default `bench.py --gpus 2` line: configs[1] data-parallel (fused all-reduce + Adam over peer memory) and configs[4]
    sharded by structure are measured on both ranks, configs[0] / [2] on rank 0."""
    line = _torchrun(29631, "--molecules", "20000", "--steps", "2", "--warmup", "3", "--cells", "256", "--ensemble-batch", "64", timeout=600)
    ow = line["other_workloads"]
    assert set(ow) == {"configs[0] ase96", "configs[1] train", "configs[2] md", "configs[4] ensemble", "large-system md (slab NVT)"}
    for k, v in ow.items():
        assert "error" not in v, (k, v)
    assert ow["configs[1] train"]["n_gpus"] == 2 and ow["configs[1] train"]["config"]["peer_barrier_status"] == 0
    assert ow["configs[4] ensemble"]["n_gpus"] == 2 and ow["configs[4] ensemble"]["config"]["cells"] == 256


@needs_two
def test_fused_allreduce_adam_matches_nccl_and_keeps_replicas_identical():
    line = _torchrun(29621, "--workload", "train", "--steps", "20", "--warmup", "3", "--no-cpu-baseline", "--verify")
    v = line["verify"]
    assert v["replicas_bitwise_identical"]
    assert v["max_abs_diff_theta32_fused_vs_nccl"] <= 1e-6 and abs(v["loss_fused"] - v["loss_nccl"]) <= 1e-6 * abs(v["loss_nccl"])


@needs_two
@pytest.mark.parametrize("cadence", [10, 0])
def test_slab_nvt_matches_single_gpu_trajectory(cadence):
    """north_star "Large-system MD": 50 NVT steps of a 60,000-atom water box on two GPUs (atoms migrate between the slabs
    at the list rebuilds) against the single-GPU loop from the same state.  The integrators are the same kernels and the
    chain sees the same global kinetic energy; the forces of the two runs differ by fp32 rounding of differently ordered
    sums (dF ~ 1e-6 eV/A on a hydrogen of 1 amu), which t = 50 x 0.049 time units turn into at most dF t^2 / 2m = 3e-6 A and
    dF t = 2.5e-6 in the momenta (measured: 2.6e-6 A, 3.4e-6).  cadence 0 = the reference's skin criterion, reduced over
    the ranks."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29641 + cadence), os.path.join(ROOT, "tests", "_slab_md_worker.py"), "20000", "50", str(cadence)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["counters"]["steps"] == 50
    if cadence:
        assert r["counters"]["rebuilds"] == 4 == r["single_rebuilds"]
    assert r["max_displacement_A"] > 0.2                            # the atoms really moved
    assert r["max_position_deviation_A"] <= 1e-5, r
    assert r["max_momentum_deviation"] <= 1e-5, r
