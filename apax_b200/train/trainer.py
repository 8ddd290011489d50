"""Mirror of the training step of apax/train/trainer.py (`make_step_fns`, :296-372; `calc_loss`, :253-263) and of the
train state it updates (flax TrainState created at apax/train/checkpoints.py:32-45), for the GMNN model.

Everything numerical happens in libapax_b200.so (`apax_train_loss_and_grad`, `apax_train_adam_step`, csrc/train.cu);
this file owns device buffers (torch = plumbing), flattens the reference's padded batch
(apax/data/input_pipeline.py:423-470) into the concatenated layout of the C ABI, and, when torch.distributed is
initialised, sums the gradient blocks over ranks (the reference shards the batch over a device mesh and lets XLA
insert the all-reduce, trainer.py:101-106, input_pipeline.py:296-300).  No oracle, no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from .. import _lib
from .._lib import check, lib
from ..optimizer.get_optimizer import AdamLinear
from ..runtime import Workspace, _ptr, _require_cuda, _stream_ptr, default_config
from .loss import LossCollection

_NAMES32 = ("emb", "w0", "b0", "w1", "b1", "w2", "b2")


class TrainState:
    """params + optimizer state on the device (flax.training.train_state.TrainState: step, params, tx, opt_state)."""

    def __init__(self, params: dict, tx: AdamLinear, config: Optional[dict] = None):
        dev = _require_cuda()
        tree = params.get("params", params)
        tree = tree.get("energy_model", tree)
        emb = np.asarray(tree["representation"]["radial_fn"]["atomic_type_embedding"], dtype=np.float32)
        ro = tree["readout"]
        if len(ro) != 3:
            raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "TrainState", "readout must be [256, 256] + output layer")
        cfg = default_config(n_out=1)
        if config:
            cfg.update(config)
        cfg["n_species"] = int(emb.shape[0])
        self.config = cfg
        c = _lib.GmConfig(n_species=cfg["n_species"], n_radial=cfg["n_radial"], n_basis=cfg["n_basis"], n_contr=cfg["n_contr"],
                          n_hidden1=cfg["nn"][0], n_hidden2=cfg["nn"][1], n_out=1, use_ntk=int(cfg["use_ntk"]),
                          use_embed_norm=int(cfg["use_embed_norm"]), mask_atoms=int(cfg["mask_atoms"]), r_min=cfg["r_min"],
                          r_max=cfg["r_max"], basis_kind=int(cfg.get("basis_kind", 0)))
        self.plan = C.c_void_p()
        check(lib().apax_train_plan_create(C.byref(c), C.byref(self.plan)), "apax_train_plan_create")
        o32 = (C.c_int64 * 8)()
        o64 = (C.c_int64 * 3)()
        check(lib().apax_train_param_layout(self.plan, C.cast(o32, C.c_void_p), C.cast(o64, C.c_void_p)), "apax_train_param_layout")
        self.off32, self.off64 = list(o32), list(o64)
        S = cfg["n_species"]
        self.shapes32 = {"emb": (S, S, cfg["n_radial"], cfg["n_basis"]), "w0": (360, 256), "b0": (256,), "w1": (256, 256),
                         "b1": (256,), "w2": (256, 1), "b2": (1,)}
        host32 = np.zeros(self.off32[7], dtype=np.float32)
        arrays = {"emb": emb}
        for li in range(3):
            arrays[f"w{li}"] = np.asarray(ro[f"dense_{li}"]["w"], dtype=np.float32)
            arrays[f"b{li}"] = np.asarray(ro[f"dense_{li}"]["b"], dtype=np.float32)
        for k, name in enumerate(_NAMES32):
            a = arrays[name]
            if tuple(a.shape) != self.shapes32[name]:
                raise _lib.ApaxB200Error(_lib.ERR_INVALID_ARG, "TrainState", f"{name} has shape {a.shape}")
            host32[self.off32[k]: self.off32[k] + a.size] = a.reshape(-1)
        ss = tree["scale_shift"]
        host64 = np.concatenate([np.broadcast_to(np.asarray(ss["scale_per_element"], dtype=np.float64).reshape(-1), (S,)),
                                 np.broadcast_to(np.asarray(ss["shift_per_element"], dtype=np.float64).reshape(-1), (S,))])
        self.theta32 = torch.as_tensor(host32).to(dev)
        self.theta64 = torch.as_tensor(host64).to(dev)
        self.grad32 = torch.zeros_like(self.theta32)
        # gradient block of scale/shift; the loss of the step sits in the tail so that one all-reduce carries both
        self.grad64 = torch.zeros(2 * S + 8, dtype=torch.float64, device=dev)
        self.loss = self.grad64[2 * S: 2 * S + 1]
        self.m32, self.v32 = torch.zeros_like(self.theta32), torch.zeros_like(self.theta32)
        self.m64, self.v64 = torch.zeros_like(self.theta64), torch.zeros_like(self.theta64)
        self.count = torch.zeros(8, dtype=torch.int64, device=dev)   # [0] = optax count, rest: device scratch
        self.tx = tx
        self._opt_c = tx.to_c()

    @property
    def step(self) -> int:
        return int(self.count[0].item())

    def _tree(self, t32: torch.Tensor, t64: torch.Tensor) -> dict:
        h32, h64 = t32.cpu().numpy(), t64.cpu().numpy()
        S = self.config["n_species"]
        arr = {n: h32[self.off32[k]: self.off32[k] + int(np.prod(self.shapes32[n]))].reshape(self.shapes32[n]).copy()
               for k, n in enumerate(_NAMES32)}
        return {"params": {"energy_model": {
            "representation": {"radial_fn": {"atomic_type_embedding": arr["emb"]}},
            "readout": {f"dense_{i}": {"w": arr[f"w{i}"], "b": arr[f"b{i}"]} for i in range(3)},
            "scale_shift": {"scale_per_element": h64[:S].reshape(S, 1).copy(), "shift_per_element": h64[S:2 * S].reshape(S, 1).copy()},
        }}}

    @property
    def params(self) -> dict:
        """The parameter tree in apax's layout (host copy)."""
        return self._tree(self.theta32, self.theta64)

    @property
    def grads(self) -> dict:
        """The gradient of the last loss evaluation, same tree layout (host copy)."""
        return self._tree(self.grad32, self.grad64[: 2 * self.config["n_species"]])

    def apply_gradients(self) -> "TrainState":
        """TrainState.apply_gradients with the gradient blocks already on the device (in place)."""
        check(lib().apax_train_adam_step(_stream_ptr(), self.plan, C.byref(self._opt_c), _ptr(self.theta32), _ptr(self.theta64),
                                         _ptr(self.grad32), _ptr(self.grad64), _ptr(self.m32), _ptr(self.v32), _ptr(self.m64),
                                         _ptr(self.v64), _ptr(self.count)), "apax_train_adam_step")
        return self

    def __del__(self):
        p = getattr(self, "plan", None)
        if p is not None and p.value:
            try:
                lib().apax_train_plan_destroy(p)
            except Exception:
                pass
            self.plan = None


def flatten_batch(inputs: dict, labels: Optional[dict] = None) -> dict:
    """The reference's padded batch -> concatenated host arrays of the C ABI.  Padding atoms (Z = 0) and padding
    entries ((0,0) pairs) stay in place; periodic structures carry fractional positions (preprocessing.py:47-54),
    which become Cartesian here (cart = box . frac, the transform of apax/layers/distances.py:12-14)."""
    pos = np.asarray(inputs["positions"], dtype=np.float64)
    B, nmax = pos.shape[0], pos.shape[1]
    box = np.asarray(inputs["box"], dtype=np.float64).reshape(B, 3, 3) if np.asarray(inputs["box"]).size == B * 9 else None
    if box is not None and np.any(np.abs(box) > 1e-6):
        periodic = np.any(np.abs(box) > 1e-6, axis=(1, 2))
        cart = np.einsum("bij,baj->bai", box, pos)
        pos = np.where(periodic[:, None, None], cart, pos)
    idx = np.asarray(inputs["idx"]).astype(np.int64)
    pmax = idx.shape[2]
    gidx = idx + (np.arange(B) * nmax)[:, None, None]
    out = {
        "R": np.ascontiguousarray(pos.reshape(B * nmax, 3)),
        "Z": np.ascontiguousarray(np.asarray(inputs["numbers"]).reshape(B * nmax).astype(np.int32)),
        "idx": np.ascontiguousarray(gidx.transpose(1, 0, 2).reshape(2, B * pmax).astype(np.int32)),
        "offsets": np.ascontiguousarray(np.asarray(inputs["offsets"], dtype=np.float64).reshape(B * pmax, 3)),
        "struct_ptr": (np.arange(B + 1) * nmax).astype(np.int32),
        "pair_ptr": (np.arange(B + 1) * pmax).astype(np.int32),
        "n_real": np.asarray(inputs["n_atoms"]).astype(np.int32).reshape(B),
        "B": B, "nmax": nmax, "pmax": pmax,
    }
    if labels is not None:
        out["E_label"] = np.ascontiguousarray(np.asarray(labels["energy"], dtype=np.float64).reshape(B))
        out["F_label"] = np.ascontiguousarray(np.asarray(labels["forces"], dtype=np.float64).reshape(B * nmax, 3))
    return out


def shard_batch(inputs: dict, labels: dict, rank: int, world: int):
    """This rank's contiguous block of the batch axis -- what sharding the batch over the ("data",) device mesh does
    in the reference (apax/train/trainer.py:101-106, apax/data/input_pipeline.py:296-300).  The batch size must be
    divisible by the number of ranks, as there."""
    B = np.asarray(inputs["positions"]).shape[0]
    if B % world:
        raise ValueError(f"batch size {B} is not divisible by {world} ranks")
    sl = slice(rank * (B // world), (rank + 1) * (B // world))
    return {k: np.asarray(v)[sl] for k, v in inputs.items()}, {k: np.asarray(v)[sl] for k, v in labels.items()}


def allreduce_sum_(dist, tensors) -> None:
    """Sum the per-rank loss / gradient blocks in place.  With batch_divisor = global batch size every rank's gradient
    is its share of the gradient of the global-batch mean loss, so a plain sum (no division) is the reference's result."""
    for t in tensors:
        dist.all_reduce(t)


_KEYS = (("R", torch.float64), ("Z", torch.int32), ("idx", torch.int32), ("offsets", torch.float64),
         ("struct_ptr", torch.int32), ("pair_ptr", torch.int32), ("n_real", torch.int32), ("E_label", torch.float64),
         ("F_label", torch.float64))


class DeviceBatch:
    """Static device buffers for batches of one shape (B, Nmax, Pmax) + pinned host staging."""

    def __init__(self, B: int, nmax: int, pmax: int):
        dev = _require_cuda()
        self.B, self.nmax, self.pmax = B, nmax, pmax
        n, P = B * nmax, B * pmax
        shapes = {"R": (n, 3), "Z": (n,), "idx": (2, P), "offsets": (P, 3), "struct_ptr": (B + 1,), "pair_ptr": (B + 1,),
                  "n_real": (B,), "E_label": (B,), "F_label": (n, 3)}
        self.dev = {k: torch.empty(shapes[k], dtype=dt, device=dev) for k, dt in _KEYS}
        self.pin = {k: torch.empty(shapes[k], dtype=dt).pin_memory() for k, dt in _KEYS}
        self.energy = torch.zeros(B, dtype=torch.float64, device=dev)
        self.forces = torch.zeros((n, 3), dtype=torch.float64, device=dev)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.pin.values())

    def upload(self, flat: dict) -> None:
        for k, _ in _KEYS:
            self.pin[k].numpy()[...] = flat[k]
            self.dev[k].copy_(self.pin[k], non_blocking=True)


class PeerGradientBlocks:
    """Gradient blocks of this rank in CUDA-IPC memory mapped by every peer (one allocation: flag array, then two
    parities of [fp32 block | fp64 block + loss]) for `apax_train_allreduce_adam_step`.  Handles travel through
    torch.distributed's object all-gather (host plumbing); the data path afterwards is this library's kernel only."""

    def __init__(self, state: TrainState, dist):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 8:
            raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "PeerGradientBlocks", "at most 8 ranks (one NVSwitch domain)")
        n32, n64 = state.off32[7], state.grad64.numel()
        b32 = (n32 * 4 + 255) // 256 * 256
        self.parity_bytes = b32 + (n64 * 8 + 255) // 256 * 256
        total = 256 + 2 * self.parity_bytes
        own, handle = C.c_void_p(), C.create_string_buffer(64)
        check(lib().apax_peer_alloc(total, C.byref(own), C.cast(handle, C.c_void_p)), "apax_peer_alloc")
        self.own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw))
        self.mapped = []
        for r in range(self.world):
            if r == self.rank:
                self.mapped.append(own.value)
            else:
                ptr = C.c_void_p()
                hb = C.create_string_buffer(handles[r], 64)
                check(lib().apax_peer_open(C.cast(hb, C.c_void_p), C.byref(ptr)), "apax_peer_open")
                self.mapped.append(ptr.value)
        self.peers = []
        for parity in range(2):
            rec = _lib.TrainPeers()
            for r in range(self.world):
                base = self.mapped[r] + 256 + parity * self.parity_bytes
                rec.grad32[r], rec.grad64[r], rec.flags[r] = base, base + b32, self.mapped[r]
            rec.world, rec.rank = self.world, self.rank
            self.peers.append(rec)
        self.local32 = [C.c_void_p(self.peers[q].grad32[self.rank]) for q in range(2)]
        self.local64 = [C.c_void_p(self.peers[q].grad64[self.rank]) for q in range(2)]
        self.local_loss = [C.c_void_p(self.peers[q].grad64[self.rank] + 8 * 2 * state.config["n_species"]) for q in range(2)]
        self.epoch = 0
        dist.barrier()          # every rank has mapped every block before the first kernel touches one

    def close(self):
        for r, ptr in enumerate(getattr(self, "mapped", [])):
            if r != self.rank and ptr:
                lib().apax_peer_close(C.c_void_p(ptr))
        if getattr(self, "own", None) is not None and self.own.value:
            lib().apax_peer_free(self.own)
        self.mapped, self.own = [], None


class TrainStep:
    """update_step of trainer.py:332-336 for one batch shape: loss + gradients, all-reduce over ranks, Adam.

    With `graph=True` the device work of a step is captured once in a CUDA graph (~35 small kernels: the step is
    launch-latency bound) and replayed; new batches are copied into the static buffers first.
    """

    def __init__(self, state: TrainState, loss_fn: LossCollection, B: int, nmax: int, pmax: int, graph: bool = False,
                 world_size: Optional[int] = None, fused_allreduce: bool = True):
        import torch.distributed as dist

        self.state, self.loss_fn = state, loss_fn
        self.batch = DeviceBatch(B, nmax, pmax)
        self._loss_c = loss_fn.to_c()
        self.dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = (self.dist.get_world_size() if self.dist else 1) if world_size is None else world_size
        self.ws = Workspace()
        n, P = B * nmax, B * pmax
        self.nbytes = lib().apax_train_workspace_bytes(state.plan, n, P, B)
        self.buf = self.ws.get(self.nbytes)
        self.graphs = [None, None]
        self._want_graph = graph
        # data parallel: all-reduce + Adam in one kernel over NVLink peer memory (default), or NCCL + the plain Adam kernels
        self.peer = PeerGradientBlocks(state, self.dist) if (self.dist and fused_allreduce) else None
        self.loss = torch.zeros(1, dtype=torch.float64, device=state.theta32.device) if self.peer else state.loss

    def _grad(self, parity: int = 0) -> None:
        st, b = self.state, self.batch
        d = b.dev
        n, P = b.B * b.nmax, b.B * b.pmax
        if self.peer:
            g32, g64, loss = self.peer.local32[parity], self.peer.local64[parity], self.peer.local_loss[parity]
        else:
            g32, g64, loss = _ptr(st.grad32), _ptr(st.grad64), _ptr(st.loss)
        check(lib().apax_train_loss_and_grad(
            _stream_ptr(), st.plan, _ptr(st.theta32), _ptr(st.theta64), _ptr(d["R"]), _ptr(d["Z"]), _ptr(d["idx"]),
            _ptr(d["offsets"]), _ptr(d["struct_ptr"]), _ptr(d["pair_ptr"]), _ptr(d["n_real"]), b.B, n, P, _ptr(d["E_label"]),
            _ptr(d["F_label"]), C.byref(self._loss_c), float(b.B * self.world), loss, _ptr(b.energy), _ptr(b.forces),
            g32, g64, C.c_void_p(Workspace.aligned_ptr(self.buf)), C.c_size_t(self.nbytes)),
            "apax_train_loss_and_grad")

    def _reduce_and_update(self, parity: int = 0) -> None:
        st = self.state
        if self.peer:
            self.peer.epoch += 1
            check(lib().apax_train_allreduce_adam_step(
                _stream_ptr(), st.plan, C.byref(st._opt_c), C.byref(self.peer.peers[parity]), self.peer.epoch, _ptr(st.theta32),
                _ptr(st.theta64), _ptr(st.m32), _ptr(st.v32), _ptr(st.m64), _ptr(st.v64), _ptr(st.count), _ptr(self.loss)),
                "apax_train_allreduce_adam_step")
            return
        if self.dist:   # gradients of the global-batch mean loss: per-rank parts add up (batch_divisor is global)
            allreduce_sum_(self.dist, (st.grad32, st.grad64))
        st.apply_gradients()

    def _device_step(self) -> None:
        parity = (self.peer.epoch & 1) if self.peer else 0
        self._grad(parity)
        self._reduce_and_update(parity)

    def __call__(self, inputs: dict, labels: dict):
        """Returns (loss tensor [1] f64, predictions {"energy": [B], "forces": [B,Nmax,3]} device tensors)."""
        b = self.batch
        b.upload(flatten_batch(inputs, labels))
        self.run_resident()
        return self.loss, {"energy": b.energy, "forces": b.forces.view(b.B, b.nmax, 3)}

    def evaluate(self, inputs: dict, labels: dict):
        """calc_loss without the update (the reference's val_step): returns (loss [1] f64, predictions), summed over ranks."""
        b = self.batch
        b.upload(flatten_batch(inputs, labels))
        parity = (self.peer.epoch & 1) if self.peer else 0
        self._grad(parity)
        if self.peer:   # this rank's share of the loss sits behind its fp64 gradient block
            loss = self.rt_view_loss(parity).clone()
        else:
            loss = self.state.loss.clone()
        if self.dist:
            self.dist.all_reduce(loss)
        return loss, {"energy": b.energy, "forces": b.forces.view(b.B, b.nmax, 3)}

    def rt_view_loss(self, parity: int) -> torch.Tensor:
        from ..runtime import device_view

        return device_view(self.peer.local_loss[parity].value, (1,), torch.float64)

    def run_resident(self) -> None:
        """One step on the batch already in the device buffers."""
        if not self._want_graph:
            self._device_step()
            return
        # Single process: the whole step (loss+gradients, Adam) is one graph.  Data parallel: only the gradient kernels are
        # captured (one graph per gradient-block parity); the fused all-reduce + Adam kernel (or NCCL + Adam) follows
        # eagerly on the same stream.
        parity = (self.peer.epoch & 1) if self.peer else 0
        body = (lambda: self._grad(parity)) if self.dist else self._device_step
        if self.graphs[parity] is None:
            st = self.state
            saved = (st.theta32, st.theta64, st.m32, st.v32, st.m64, st.v64, st.count)
            keep = [t.clone() for t in saved]
            side = torch.cuda.Stream()   # warm up outside the capture (module load, first touch), then restore the state
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for t, k in zip(saved, keep):
                t.copy_(k)
            self.graphs[parity] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graphs[parity]):
                body()
        self.graphs[parity].replay()
        if self.dist:
            self._reduce_and_update(parity)


class TableModel:
    """Energy + forces on the table-driven kernels (apax_train_energy_forces): the evaluation path for model statics the
    specialised apax_gm_* kernels do not cover (Bessel / exponential-spacing bases, n_basis <= 16)."""

    def __init__(self, params: dict, config: dict):
        frozen = AdamLinear(emb_lr=0.0, nn_lr=0.0, scale_lr=0.0, shift_lr=0.0, gradient_clipping=1.0, transition_steps=1)
        self.state = TrainState(params, frozen, config)
        self.ws = Workspace()

    def energy_forces_batched(self, R, Z, idx, offsets, struct_ptr, pair_ptr):
        """Concatenated batch as in apax_gm_energy_forces_batched -> (energy [B] f64, forces [N,3] f64) device tensors."""
        from ..runtime import to_device

        st = self.state
        R, Z, idx = to_device(R, torch.float64), to_device(Z, torch.int32), to_device(idx, torch.int32)
        off = None if offsets is None else to_device(offsets, torch.float64)
        sp, pp = to_device(struct_ptr, torch.int32), to_device(pair_ptr, torch.int32)
        n, P, B = int(R.shape[0]), int(idx.shape[1]), int(sp.shape[0]) - 1
        energy = torch.empty(B, dtype=torch.float64, device=R.device)
        forces = torch.empty((n, 3), dtype=torch.float64, device=R.device)
        nbytes = lib().apax_train_workspace_bytes(st.plan, n, P, B)
        buf = self.ws.get(nbytes)
        check(lib().apax_train_energy_forces(_stream_ptr(), st.plan, _ptr(st.theta32), _ptr(st.theta64), _ptr(R), _ptr(Z), _ptr(idx),
                                             _ptr(off), _ptr(sp), _ptr(pp), B, n, P, _ptr(energy), _ptr(forces),
                                             C.c_void_p(Workspace.aligned_ptr(buf)), C.c_size_t(nbytes)), "apax_train_energy_forces")
        return energy, forces

    def energy_forces(self, R, Z, idx, box, offsets, mode: str):
        """One structure with the call shape of EnergyDerivativeModel.apply.  Fractional positions of the `offsets` branch
        become Cartesian (box . R, apax/layers/distances.py:12-14); the min-image branch is not offered on this path."""
        from ..runtime import to_device

        if mode == "minimage":
            raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "TableModel", "min-image displacement with a non-default basis: pass a "
                                     "list with offsets (apax_nl_build_images)")
        R = to_device(R, torch.float64)
        if mode == "offsets":
            R = R @ to_device(box, torch.float64).T
        else:
            offsets = None
        n, P = int(R.shape[0]), int(np.asarray(idx.shape)[1])
        e, f = self.energy_forces_batched(R, Z, idx, offsets, np.array([0, n], dtype=np.int32), np.array([0, P], dtype=np.int32))
        return e, f


_table_cache: dict = {}


def table_model(params: dict, config: dict) -> TableModel:
    key = (id(params), tuple(sorted((k, str(v)) for k, v in config.items())))
    hit = _table_cache.get(key)
    if hit is None or hit[0] is not params:
        _table_cache[key] = (params, TableModel(params, config))
    return _table_cache[key][1]


def make_step_fns(loss_fn: LossCollection, Metrics=None, model=None, is_ensemble: bool = False, graph: bool = False):
    """(train_step, val_step) with the reference's calling convention (trainer.py:346-372):
    train_step((state, batch_metrics), (inputs, labels)) -> ((state, batch_metrics), loss)."""
    if is_ensemble:
        raise NotImplementedError("full-ensemble training (vmap over train states) is outside the GMNN hot path")
    steps = {}

    def _get(state, inputs):
        pos, idx = np.asarray(inputs["positions"]), np.asarray(inputs["idx"])
        key = (id(state), pos.shape[0], pos.shape[1], idx.shape[2])
        if key not in steps:
            steps[key] = TrainStep(state, loss_fn, pos.shape[0], pos.shape[1], idx.shape[2], graph=graph)
        return steps[key]

    def train_step(carry, batch):
        state, batch_metrics = carry
        inputs, labels = batch
        loss, _pred = _get(state, inputs)(inputs, labels)
        return (state, batch_metrics), loss

    def val_step(state, batch, batch_metrics=None):
        """eval_fn of the reference (trainer.py:340-344, 360-370): loss and predictions of the batch, no update."""
        inputs, labels = batch
        loss, pred = _get(state, inputs).evaluate(inputs, labels)
        return loss, batch_metrics, pred

    return train_step, val_step
