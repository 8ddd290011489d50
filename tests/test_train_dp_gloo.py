"""Data-parallel training step (SURVEY.md s8e "Training") on CPU with gloo, world_size 2: the same batch sharding and
gradient all-reduce as the GPU path (apax_b200/train/trainer.py::shard_batch, allreduce_sum_), with the training oracle
injected as the local loss+gradient.  Every rank must end with the single-process full-batch gradients and parameters."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apax_b200 import synthetic as syn
from apax_b200.train.trainer import allreduce_sum_, shard_batch
from oracle import train_oracle as tor

NAMES = ("emb", "w0", "b0", "w1", "b1", "w2", "b2", "scale", "shift")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    inputs, labels = tor.ethanol_training_batch(4, seed=8, pad_atoms=1)
    params = syn.gmnn_params(seed=21, random_bias=True, random_scale_shift=True)
    inp, lab = shard_batch(inputs, labels, rank, world)
    r = tor.loss_and_grads(params, inp, lab, batch_divisor=4)          # divisor = GLOBAL batch
    tensors = [torch.as_tensor(r["grads"][k].copy()) for k in NAMES] + [torch.as_tensor([r["loss"]], dtype=torch.float64)]
    allreduce_sum_(dist, tensors)
    grads = {k: t.numpy() for k, t in zip(NAMES, tensors)}
    theta, _ = tor.adam_update(tor.params_to_flat(params), grads, tor.AdamState(), tor.OptConfig())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=tensors[-1].numpy(), **{"g_" + k: grads[k] for k in NAMES},
             **{"t_" + k: theta[k] for k in NAMES})
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_full_batch_step(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    inputs, labels = tor.ethanol_training_batch(4, seed=8, pad_atoms=1)
    params = syn.gmnn_params(seed=21, random_bias=True, random_scale_shift=True)
    full = tor.loss_and_grads(params, inputs, labels)
    theta, _ = tor.adam_update(tor.params_to_flat(params), full["grads"], tor.AdamState(), tor.OptConfig())
    r0, r1 = (dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world))
    assert abs(float(r0["loss"][0]) - full["loss"]) <= 1e-12 * abs(full["loss"])
    for k in NAMES:
        assert np.array_equal(r0["g_" + k], r1["g_" + k]) and np.array_equal(r0["t_" + k], r1["t_" + k]), k
        g = full["grads"][k]
        assert np.abs(r0["g_" + k] - g).max() <= 2e-6 * (np.abs(g).max() + 1e-30), k
        moved = np.abs(g) > 1e-4 * np.abs(g).max()        # Adam's first step is sign(g) * lr: compare where g is not noise
        assert np.allclose(r0["t_" + k][moved], theta[k][moved], rtol=1e-5, atol=1e-7), k
