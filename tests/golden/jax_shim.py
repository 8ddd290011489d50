"""Torch-backed stand-ins for the few pieces of `jax`, `flax.linen` and `ase` that apax's
hot-path source files import, so that those files can be executed UNMODIFIED from
/root/reference in a container that has no jax/flax wheels.

This is fixture-generation tooling, used only by `make_golden.py` (run by hand in the build
container; /root/reference does not exist on the GPU box).  It is not an oracle and not product
code.  What it reproduces: the reference's own Python control flow, index conventions, einsum
strings, parameter-tree scoping and custom-JVP rule, with torch-CPU kernels underneath and
torch.autograd in place of jax.grad.  What it does not reproduce: XLA's numerics.
"""
from __future__ import annotations

import dataclasses as _dc
import functools
import sys
import types

import numpy as np
import torch

_T = torch.Tensor


# --------------------------------------------------------------------------- tensor sugar
def _astype(self, dt):
    return self.to(_to_torch_dtype(dt))


class _AtIndexer:
    def __init__(self, t):
        self.t = t

    def __getitem__(self, index):
        return _AtSetter(self.t, index)


class _AtSetter:
    def __init__(self, t, index):
        self.t, self.index = t, index

    def _idx(self):
        idx = self.index
        if isinstance(idx, tuple):
            idx = tuple(i.long() if isinstance(i, _T) and not i.dtype == torch.bool else i for i in idx)
        elif isinstance(idx, _T) and idx.dtype != torch.bool:
            idx = idx.long()
        return idx

    def set(self, values):
        out = self.t.clone()
        out[self._idx()] = values if not isinstance(values, _T) else values.to(out.dtype)
        return out

    def add(self, values):
        out = self.t.clone()
        idx = self._idx()
        out[idx] = out[idx] + (values if not isinstance(values, _T) else values.to(out.dtype))
        return out

    def multiply(self, values):
        out = self.t.clone()
        idx = self._idx()
        out[idx] = out[idx] * (values if not isinstance(values, _T) else values.to(out.dtype))
        return out


_orig_getitem = torch.Tensor.__getitem__


def _clamping_getitem(self, index):
    # JAX gathers clamp out-of-range integer-array indices (numpy/torch raise instead)
    if isinstance(index, _T) and index.dtype in (torch.int32, torch.int64, torch.int16) and self.dim() > 0:
        index = index.long().clamp(-self.shape[0], self.shape[0] - 1)
    elif isinstance(index, tuple):
        new, dim = [], 0
        for it in index:
            if it is Ellipsis:
                new.append(it)
                break
            if isinstance(it, _T) and it.dtype in (torch.int32, torch.int64, torch.int16):
                it = it.long().clamp(-self.shape[dim], self.shape[dim] - 1)
            new.append(it)
            if it is not None:
                dim += 1
        index = tuple(new) + tuple(index[len(new):])
    return _orig_getitem(self, index)


torch.Tensor.__getitem__ = _clamping_getitem
torch.Tensor.astype = _astype
torch.Tensor.at = property(lambda self: _AtIndexer(self))


class _DType:
    """Callable dtype object: jnp.float32(x) casts, and it converts to a torch dtype."""

    def __init__(self, tdt, npdt):
        self.tdt, self.npdt = tdt, npdt

    def __call__(self, x):
        if isinstance(x, _T):
            return x.to(self.tdt)
        return torch.as_tensor(np.asarray(x), dtype=self.tdt)

    def __repr__(self):
        return f"shim.{self.tdt}"

    def __eq__(self, other):
        if isinstance(other, _DType):
            return self.tdt == other.tdt
        if isinstance(other, torch.dtype):
            return self.tdt == other
        return NotImplemented

    __req__ = __eq__

    def __hash__(self):
        return hash(self.tdt)


float16 = _DType(torch.float16, np.float16)
float32 = _DType(torch.float32, np.float32)
float64 = _DType(torch.float64, np.float64)
int16 = _DType(torch.int16, np.int16)
int8 = _DType(torch.int8, np.int8)
int32 = _DType(torch.int32, np.int32)
int64 = _DType(torch.int64, np.int64)
uint8 = _DType(torch.uint8, np.uint8)


def _to_torch_dtype(dt):
    if dt is None:
        return None
    if isinstance(dt, _DType):
        return dt.tdt
    if isinstance(dt, torch.dtype):
        return dt
    if dt is float:
        return torch.float64
    if dt is int:
        return torch.int64
    if dt is bool:
        return torch.bool
    return {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64,
            np.dtype("int32"): torch.int32, np.dtype("int64"): torch.int64,
            np.dtype("int16"): torch.int16, np.dtype("uint8"): torch.uint8,
            np.dtype("bool"): torch.bool}[np.dtype(dt)]


def _t(x, dtype=None):
    tdt = _to_torch_dtype(dtype)
    if isinstance(x, _T):
        return x if tdt is None else x.to(tdt)
    if isinstance(x, (list, tuple)) and len(x) and any(isinstance(e, _T) for e in x):
        return torch.stack([_t(e, dtype) for e in x])
    a = np.asarray(x)
    if a.dtype == np.float64 and tdt is None and not isinstance(x, np.ndarray):
        tdt = torch.float64  # x64 mode: python floats are float64
    return torch.as_tensor(a, dtype=tdt)


# --------------------------------------------------------------------------- jax.numpy
jnp = types.ModuleType("jax.numpy")
jnp.float16, jnp.float32, jnp.float64, jnp.float_ = float16, float32, float64, float64
jnp.int32, jnp.int64, jnp.uint8, jnp.int16, jnp.int8 = int32, int64, uint8, int16, int8
jnp.bool_ = _DType(torch.bool, np.bool_)
jnp.ndarray = _T
jnp.pi = np.pi
jnp.newaxis = None
jnp.array = lambda x, dtype=None: _t(x, dtype).clone() if isinstance(x, _T) else _t(x, dtype)
jnp.asarray = lambda x, dtype=None: _t(x, dtype)
jnp.arange = lambda *a, dtype=None: torch.arange(*a, dtype=_to_torch_dtype(dtype))
jnp.zeros = lambda shape, dtype=None: torch.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=_to_torch_dtype(dtype) or torch.float64)
jnp.eye = lambda n, dtype=None: torch.eye(n, dtype=_to_torch_dtype(dtype) or torch.float64)
jnp.ones = lambda shape, dtype=None: torch.ones(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=_to_torch_dtype(dtype) or torch.float64)
jnp.full = lambda shape, v, dtype=None: torch.full(tuple(shape), v, dtype=_to_torch_dtype(dtype) or (torch.int64 if isinstance(v, int) else torch.float64))
jnp.zeros_like = torch.zeros_like
jnp.einsum = lambda spec, *ops: torch.einsum(spec.replace(" ", ""), *ops)
jnp.exp = lambda x: torch.exp(_t(x))
jnp.cos = torch.cos
jnp.sin = torch.sin
jnp.sinc = lambda x: torch.sinc(_t(x))
jnp.sqrt = lambda x: torch.sqrt(_t(x))
jnp.abs = torch.abs
jnp.log = lambda x: torch.log(_t(x))
jnp.sign = torch.sign
jnp.maximum = lambda a, b: torch.maximum(_t(a), _t(b))
jnp.dot = lambda a, b: torch.matmul(a, b)
jnp.matmul = torch.matmul
jnp.stack = lambda xs, axis=0: torch.stack(list(xs), dim=axis)
jnp.concatenate = lambda xs, axis=0: torch.cat(list(xs), dim=axis)
jnp.reshape = lambda x, shape: torch.reshape(x, tuple(shape))
jnp.transpose = lambda x, axes=None: x.permute(*axes) if axes is not None else x.T
jnp.broadcast_to = lambda x, shape: torch.broadcast_to(_t(x), tuple(shape))
jnp.cumsum = lambda x, axis=None: torch.cumsum(x.reshape(-1) if axis is None else x, dim=0 if axis is None else axis)
jnp.any = lambda x: torch.any(x)
jnp.all = lambda x: torch.all(x)
jnp.isscalar = lambda x: isinstance(x, (int, float)) or (isinstance(x, _T) and x.dim() == 0 and False)
jnp.mod = lambda a, b: torch.remainder(_t(a), b if not isinstance(b, _T) else b)


def _where(c, a, b):
    a_is, b_is = isinstance(a, _T), isinstance(b, _T)
    if a_is and not b_is:
        b = torch.as_tensor(b, dtype=a.dtype)
    elif b_is and not a_is:
        a = torch.as_tensor(a, dtype=b.dtype)
    elif not a_is and not b_is:
        a, b = _t(a), _t(b)
    return torch.where(c, a, b)


jnp.where = _where
jnp.integer = "integer"
jnp.complexfloating = "complexfloating"
jnp.floating = "floating"


def _issubdtype(dt, kind):
    tdt = _to_torch_dtype(dt)
    if kind == "integer":
        return tdt in (torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8)
    if kind == "complexfloating":
        return tdt in (torch.complex64, torch.complex128)
    return tdt in (torch.float16, torch.float32, torch.float64)


jnp.issubdtype = _issubdtype


def _clip(x, min=None, max=None):
    return torch.clamp(x, min=min, max=max)


jnp.clip = _clip


def _sum(x, axis=None, dtype=None, keepdims=False):
    tdt = _to_torch_dtype(dtype)
    if axis is None:
        return torch.sum(x, dtype=tdt) if not keepdims else torch.sum(x, dim=tuple(range(x.dim())), keepdim=True, dtype=tdt)
    return torch.sum(x, dim=axis, keepdim=keepdims, dtype=tdt)


def _mean(x, axis=None):
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis)


def _max(x, axis=None):
    return torch.max(x) if axis is None else torch.max(x, dim=axis).values


jnp.sum, jnp.mean, jnp.max = _sum, _mean, _max
_linalg = types.ModuleType("jax.numpy.linalg")
_linalg.inv = torch.linalg.inv
_linalg.norm = lambda x, ord=None, axis=None, keepdims=False: torch.linalg.norm(x, ord=ord, dim=axis, keepdim=keepdims)
jnp.linalg = _linalg


# --------------------------------------------------------------------------- jax core
def _needs_grad(a):
    while torch._C._functorch.is_batchedtensor(a):
        a = torch._C._functorch.get_unwrapped(a)
    return a.requires_grad


class custom_jvp:
    """Generic custom_jvp: the primal is evaluated on detached inputs and the user's own JVP
    rule is evaluated on zero-valued tangents that carry autograd identity w.r.t. the inputs,
    so reverse mode sees exactly the transpose of the registered (linear) rule."""

    def __init__(self, fn):
        self.fn, self.rule = fn, None
        functools.update_wrapper(self, fn)

    def defjvp(self, rule):
        self.rule = rule
        return rule

    def __call__(self, *args):
        targs = [a if isinstance(a, _T) else _t(a) for a in args]
        prim = [a.detach() for a in targs]
        if not any(_needs_grad(a) for a in targs):
            return self.fn(*prim)  # also terminates the rule's own recursive transform() calls
        tang = [a - a.detach() for a in targs]
        out, tout = self.rule(tuple(prim), tuple(tang))
        return out.detach() + tout


def jit(fn=None, static_argnums=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def vmap(fn, in_axes=0, out_axes=0):
    return torch.vmap(fn, in_dims=in_axes, out_dims=out_axes)


HIGHER_ORDER = False


def value_and_grad(fn, has_aux=False):
    def wrapped(x, *args, **kwargs):
        x = _t(x).detach().clone().requires_grad_(True)
        out = fn(x, *args, **kwargs)
        val = out[0] if has_aux else out
        # HIGHER_ORDER: keep the graph of the gradient itself (the training step differentiates the
        # forces w.r.t. the weights, apax/train/trainer.py:332)
        (g,) = torch.autograd.grad(val, x, create_graph=HIGHER_ORDER)
        return out, g

    return wrapped


def grad(fn, has_aux=False):
    vg = value_and_grad(fn, has_aux)

    def wrapped(*a, **k):
        return vg(*a, **k)[1]

    return wrapped


def jacrev(fn):
    def wrapped(x, *args, **kwargs):
        x = _t(x).detach().clone().requires_grad_(True)
        out = fn(x, *args, **kwargs)
        rows = [torch.autograd.grad(out[k], x, retain_graph=True)[0] for k in range(out.shape[0])]
        return torch.stack(rows)

    return wrapped


def eval_shape(fn, *args, **kwargs):
    conc = [torch.zeros(a.shape, dtype=_to_torch_dtype(a.dtype)) if isinstance(a, ShapedArray) else a for a in args]
    return fn(*conc, **kwargs)


class ShapedArray:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), dtype


def _segment_sum(data, segment_ids, num_segments=None):
    seg = segment_ids.long()
    valid = (seg >= 0) & (seg < num_segments)
    seg_c = torch.where(valid, seg, torch.zeros_like(seg))
    w = valid.to(data.dtype).reshape((-1,) + (1,) * (data.dim() - 1))
    out = torch.zeros((num_segments,) + tuple(data.shape[1:]), dtype=data.dtype)
    return out.index_add(0, seg_c, data * w)


class _Lax(types.ModuleType):
    @staticmethod
    def stop_gradient(x):
        return x.detach() if isinstance(x, _T) else x

    @staticmethod
    def scan(f, init, xs, length=None, unroll=1):
        carry, ys = init, []
        for i in range(len(xs)):
            carry, y = f(carry, xs[i])
            ys.append(y)
        if ys and isinstance(ys[0], _T):
            ys = torch.stack(ys)
        return carry, ys

    @staticmethod
    def cond(pred, *args):
        # lax.cond(pred, operand_true..., true_fn, operand_false, false_fn) legacy 5-arg form
        if len(args) == 4:
            t_op, t_fn, f_op, f_fn = args
            return t_fn(t_op) if bool(pred) else f_fn(f_op)
        t_fn, f_fn, *ops = args
        return t_fn(*ops) if bool(pred) else f_fn(*ops)


def install(reference_root="/root/reference"):
    """Register the shim modules and bare `apax` package stubs (so that apax/__init__.py and the
    lazy-loader inits, which pull in the whole dependency closure, are not executed)."""
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.custom_jvp, jax.jit, jax.vmap = custom_jvp, jit, vmap
    jax.value_and_grad, jax.grad, jax.jacrev = value_and_grad, grad, jacrev
    jax.eval_shape = eval_shape
    jax.Array = _T

    class ShapeDtypeStruct:  # eval_shape returns concrete tensors in this shim
        pass

    jax.ShapeDtypeStruct = ShapeDtypeStruct
    jscipy = types.ModuleType("jax.scipy")
    jspecial = types.ModuleType("jax.scipy.special")
    jspecial.gammaln = torch.lgamma
    jscipy.special = jspecial
    jax.scipy = jscipy
    jax.lax = _Lax("jax.lax")
    jax.ensure_compile_time_eval = lambda: __import__("contextlib").nullcontext()
    tree_util = types.ModuleType("jax.tree_util")
    tree_util.register_pytree_node = lambda *a, **k: None
    def _tree_map(f, t, *rest):
        if isinstance(t, dict):
            return {k: _tree_map(f, v, *[r[k] for r in rest]) for k, v in t.items()}
        if isinstance(t, (list, tuple)):
            return type(t)(_tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(t))
        return f(t, *rest)

    def _tree_leaves(t):
        if isinstance(t, dict):
            return [l for v in t.values() for l in _tree_leaves(v)]
        if isinstance(t, (list, tuple)):
            return [l for v in t for l in _tree_leaves(v)]
        return [t]

    tree_util.tree_map = _tree_map
    tree_util.tree_reduce = lambda f, t, init: functools.reduce(f, _tree_leaves(t), init)
    tree_util.tree_flatten = lambda t: (_tree_leaves(t), None)
    tree_util.tree_unflatten = lambda treedef, leaves: leaves[0]
    jax.tree_util = tree_util
    ops = types.ModuleType("jax.ops")
    ops.segment_sum = _segment_sum
    jax.ops = ops
    jnn = types.ModuleType("jax.nn")
    jnn.swish = lambda x: x * torch.sigmoid(x)
    jnn.silu = jnn.swish
    jnn.softplus = lambda x: torch.nn.functional.softplus(_t(x))
    inits = types.ModuleType("jax.nn.initializers")
    inits.Initializer = object
    jnn.initializers = inits
    jax.nn = jnn
    jrandom = types.ModuleType("jax.random")
    jax.random = jrandom
    jcore = types.ModuleType("jax.core")
    jcore.ShapedArray = ShapedArray
    jax.core = jcore
    jsrc = types.ModuleType("jax._src")
    jsrc_dtypes = types.ModuleType("jax._src.dtypes")
    jsrc.dtypes = jsrc_dtypes
    jax._src = jsrc
    mods = {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": _linalg, "jax.lax": jax.lax,
            "jax.tree_util": tree_util, "jax.ops": ops, "jax.nn": jnn, "jax.nn.initializers": inits,
            "jax.random": jrandom, "jax.core": jcore, "jax._src": jsrc, "jax._src.dtypes": jsrc_dtypes,
            "jax.scipy": jscipy, "jax.scipy.special": jspecial}

    # ---- flax.linen
    flax = types.ModuleType("flax")
    linen = types.ModuleType("flax.linen")
    linen.Module = Module
    linen.compact = lambda f: f
    linen.Sequential = Sequential
    li = types.ModuleType("flax.linen.initializers")
    li.normal = lambda *a, **k: ("normal", a, k)
    li.lecun_normal = lambda *a, **k: ("lecun_normal", a, k)
    li.constant = lambda *a, **k: ("constant", a, k)
    linen.initializers = li
    flax.linen = linen
    mods.update({"flax": flax, "flax.linen": linen, "flax.linen.initializers": li})

    # ---- ase (only names imported at module level by apax/utils/convert.py, layers/empirical.py)
    ase = types.ModuleType("ase")
    ase.Atoms = object
    units = types.ModuleType("ase.units")
    for k in ("Ang", "Bohr", "Hartree", "eV", "kcal", "kJ", "mol"):
        setattr(units, k, 1.0)
    ase.units = units
    data = types.ModuleType("ase.data")
    data.covalent_radii = np.full(119, 1.0)   # only the initialiser of ExponentialRepulsion reads it; parameters are passed explicitly
    ase.data = data
    mods.update({"ase": ase, "ase.units": units, "ase.data": data})

    # ---- apax package stubs with real __path__ (submodules import from the reference tree)
    for pkg in ("apax", "apax.layers", "apax.layers.descriptor", "apax.utils", "apax.utils.jax_md_reduced",
                "apax.nn", "apax.md", "apax.train"):
        m = types.ModuleType(pkg)
        m.__path__ = [reference_root + "/" + pkg.replace(".", "/")]
        mods[pkg] = m
    sys.modules.update(mods)
    import importlib

    # apax.utils.jax_md_reduced's __init__ would import simulate/smap/quantity; expose what is used
    jm = sys.modules["apax.utils.jax_md_reduced"]
    smap = types.ModuleType("apax.utils.jax_md_reduced.smap")  # pair-potential mappers: unused by GMNN
    sys.modules["apax.utils.jax_md_reduced.smap"] = smap
    jm.smap = smap
    for sub in ("util", "dataclasses", "space", "partition", "quantity", "simulate"):
        setattr(jm, sub, importlib.import_module(f"apax.utils.jax_md_reduced.{sub}"))
    return jax


# --------------------------------------------------------------------------- flax.linen.Module
_CTX: list = []  # stack of modules whose setup()/__call__ is executing


class Module:
    name: str | None = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        # flax turns every subclass into a dataclass; keep `name` (and parent) as trailing kw fields
        ann = dict(cls.__dict__.get("__annotations__", {}))
        ann.pop("name", None)
        ann["name"] = "str | None"
        cls.__annotations__ = ann
        cls.name = None
        _dc.dataclass(eq=False, repr=False)(cls)
        user_call = cls.__dict__.get("__call__")
        if user_call is not None:
            @functools.wraps(user_call)
            def wrapped(self, *a, __f=user_call, **k):
                self._ensure_setup()
                _CTX.append(self)
                try:
                    return __f(self, *a, **k)
                finally:
                    _CTX.pop()

            cls.__call__ = wrapped

    def __post_init__(self):
        object.__setattr__(self, "_parent", _CTX[-1] if _CTX else None)
        object.__setattr__(self, "_params", None)
        object.__setattr__(self, "_setup_done", False)
        if self._parent is not None and self.name is not None and self._parent._params is not None:
            object.__setattr__(self, "_params", self._parent._params.get(self.name, {}))

    # binding -----------------------------------------------------------------------------
    def _adopt(self, child, name):
        if isinstance(child, Module):
            if child._params is None:
                nm = child.name or name
                object.__setattr__(child, "_parent", self)
                object.__setattr__(child, "_params", (self._params or {}).get(nm, {}))
        elif isinstance(child, (list, tuple)):
            for i, c in enumerate(child):
                self._adopt(c, f"{name}_{i}")

    def __setattr__(self, key, value):
        if getattr(self, "_in_setup", False):
            self._adopt(value, key)
        object.__setattr__(self, key, value)

    def _ensure_setup(self):
        if self._setup_done:
            return
        object.__setattr__(self, "_setup_done", True)
        for f in _dc.fields(self):
            if f.name != "name":
                self._adopt(getattr(self, f.name), f.name)
        if hasattr(self, "setup"):
            object.__setattr__(self, "_in_setup", True)
            _CTX.append(self)
            try:
                self.setup()
            finally:
                _CTX.pop()
                object.__setattr__(self, "_in_setup", False)

    def __getattr__(self, item):
        # attributes defined in setup() are created lazily, as in flax
        if item.startswith("_"):
            raise AttributeError(item)
        if not self.__dict__.get("_setup_done", False) and self.__dict__.get("_params") is not None:
            self._ensure_setup()
            if item in self.__dict__:
                return self.__dict__[item]
        raise AttributeError(f"{type(self).__name__}.{item}")

    # parameters --------------------------------------------------------------------------
    def param(self, name, init_fn, *init_args):
        if self._params is None or name not in self._params:
            raise KeyError(f"missing parameter {name!r} in scope of {type(self).__name__}({self.name})")
        val = self._params[name]
        val = val if isinstance(val, _T) else torch.as_tensor(np.asarray(val))
        if init_args and isinstance(init_args[0], (tuple, list)):
            shp = tuple(int(s) for s in init_args[0])
            assert tuple(val.shape) == shp, (name, tuple(val.shape), shp)
        return val

    def apply(self, variables, *args, **kwargs):
        root = _dc.replace(self)  # fresh, unbound copy (children are re-adopted lazily)
        _unbind(root)
        object.__setattr__(root, "_params", variables["params"])
        return root(*args, **kwargs)


def _unbind(m):
    if isinstance(m, Module):
        object.__setattr__(m, "_params", None)
        object.__setattr__(m, "_setup_done", False)
        for f in _dc.fields(m):
            if f.name != "name":
                _unbind(getattr(m, f.name))
    elif isinstance(m, (list, tuple)):
        for c in m:
            _unbind(c)


class Sequential(Module):
    layers: list = None

    def __call__(self, x):
        for layer in self.layers:
            x = layer(x)
        return x
