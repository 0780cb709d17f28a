"""NumPy stand-ins for the slice of `jax` / `jax.numpy` the reference's hot-path files touch.

TEST INFRASTRUCTURE (see oracle/refshim/__init__.py).  Not a JAX re-implementation: arrays are
numpy arrays with jax's *immutable* surface (`x.at[i].set(v)`, augmented assignment rebinding
instead of writing in place, because the reference does `mask &= broadcast_view`), pytrees are
dict / list / tuple / None with sorted dict keys like jax, and jit is the identity.
Arithmetic runs in whatever dtype numpy picks (float64 when the inputs are float64), which is
what makes the reference-run fixtures a ground truth rather than an fp32 sample.
"""
from __future__ import annotations

import types

import numpy as np


# ----------------------------------------------------------------------------------------------
# arrays
# ----------------------------------------------------------------------------------------------
class _AtIndex:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def _new(self):
        return np.array(self._arr, copy=True).view(Arr)

    def set(self, value):
        out = self._new()
        np.ndarray.__setitem__(out, self._idx, value)
        return out

    def add(self, value):
        out = self._new()
        np.ndarray.__setitem__(out, self._idx, np.ndarray.__getitem__(out, self._idx) + value)
        return out

    def multiply(self, value):
        out = self._new()
        np.ndarray.__setitem__(out, self._idx, np.ndarray.__getitem__(out, self._idx) * value)
        return out


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self._arr, idx)


def _binary_out_of_place(name):
    op = getattr(np.ndarray, name)

    def inplace(self, other):
        return op(self, other)
    return inplace


class Arr(np.ndarray):
    """ndarray with jax.Array's functional-update surface."""

    @property
    def at(self):
        return _At(self)

    # jax arrays are immutable: `a += b` rebinds `a`
    __iadd__ = _binary_out_of_place("__add__")
    __isub__ = _binary_out_of_place("__sub__")
    __imul__ = _binary_out_of_place("__mul__")
    __itruediv__ = _binary_out_of_place("__truediv__")
    __iand__ = _binary_out_of_place("__and__")
    __ior__ = _binary_out_of_place("__or__")

    def block_until_ready(self):
        return self


def wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, Arr):
        return x.view(Arr)
    if isinstance(x, tuple):
        return tuple(wrap(v) for v in x)
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x


def _wrapped(fn):
    def call(*args, **kwargs):
        return wrap(fn(*args, **kwargs))
    call.__name__ = getattr(fn, "__name__", "fn")
    return call


class _NumpyNamespace(types.ModuleType):
    """`jax.numpy`: every numpy function, results viewed as `Arr`."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = getattr(np, name)
        if callable(obj) and not isinstance(obj, type):
            obj = _wrapped(obj)
        setattr(self, name, obj)
        return obj


def make_jnp() -> types.ModuleType:
    jnp = _NumpyNamespace("jax.numpy")
    jnp.ndarray = np.ndarray
    jnp.array = lambda x, dtype=None, copy=True: np.array(x, dtype=dtype, copy=copy).view(Arr)
    jnp.asarray = lambda x, dtype=None: np.asarray(x, dtype=dtype).view(Arr)
    jnp.finfo, jnp.iinfo = np.finfo, np.iinfo
    jnp.bool_ = np.bool_
    jnp.newaxis, jnp.pi, jnp.inf = None, np.pi, np.inf
    jnp.einsum = _wrapped(np.einsum)
    return jnp


# ----------------------------------------------------------------------------------------------
# pytrees (dict / list / tuple / None), sorted dict keys as in jax
# ----------------------------------------------------------------------------------------------
class DictKey:
    def __init__(self, key):
        self.key = key

    def __repr__(self):
        return f"['{self.key}']"


class SequenceKey:
    def __init__(self, idx):
        self.idx = idx
        self.key = idx

    def __repr__(self):
        return f"[{self.idx}]"


def _is_node(x):
    return isinstance(x, (dict, list, tuple)) or x is None


def _map(f, path, tree, rest, with_path, is_leaf):
    if is_leaf is not None and is_leaf(tree):
        return f(path, tree, *rest) if with_path else f(tree, *rest)
    if tree is None:
        return None
    if isinstance(tree, dict):
        return {k: _map(f, path + (DictKey(k),), tree[k], [r[k] for r in rest], with_path, is_leaf)
                for k in sorted(tree.keys())}
    if isinstance(tree, (list, tuple)):
        out = [_map(f, path + (SequenceKey(i),), v, [r[i] for r in rest], with_path, is_leaf)
               for i, v in enumerate(tree)]
        return type(tree)(out) if not hasattr(tree, "_fields") else type(tree)(*out)
    return f(path, tree, *rest) if with_path else f(tree, *rest)


def tree_map(f, tree, *rest, is_leaf=None):
    return _map(f, (), tree, list(rest), False, is_leaf)


def tree_map_with_path(f, tree, *rest, is_leaf=None):
    return _map(f, (), tree, list(rest), True, is_leaf)


def tree_leaves(tree, is_leaf=None):
    out = []
    tree_map(lambda x: out.append(x), tree, is_leaf=is_leaf)
    return out


class _TreeDef:
    def __init__(self, tree):
        self.skeleton = tree_map(lambda _: 0, tree)
        self.num_leaves = len(tree_leaves(tree))

    def unflatten(self, leaves):
        it = iter(leaves)
        return tree_map(lambda _: next(it), self.skeleton)


def tree_flatten(tree, is_leaf=None):
    return tree_leaves(tree, is_leaf), _TreeDef(tree)


def tree_unflatten(treedef, leaves):
    return treedef.unflatten(leaves)


# ----------------------------------------------------------------------------------------------
# random (values only need to be deterministic, not jax's threefry stream)
# ----------------------------------------------------------------------------------------------
def PRNGKey(seed):
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32).view(Arr)


def _seed_of(key):
    k = np.asarray(key, dtype=np.uint64).ravel()
    return int((k[0] << np.uint64(32)) | k[-1])


def split(key, num=2):
    ss = np.random.SeedSequence(_seed_of(key)).generate_state(2 * num, dtype=np.uint32)
    return ss.reshape(num, 2).view(Arr)


def rng_of(key):
    return np.random.default_rng(_seed_of(key))


def random_normal(key, shape=(), dtype=np.float64):
    return rng_of(key).standard_normal(tuple(shape)).astype(dtype).view(Arr)


def random_uniform(key, shape=(), dtype=np.float64, minval=0.0, maxval=1.0):
    return rng_of(key).uniform(minval, maxval, tuple(shape)).astype(dtype).view(Arr)


# ----------------------------------------------------------------------------------------------
# transforms
# ----------------------------------------------------------------------------------------------
def jit(fn=None, **_kwargs):
    if fn is None:
        return lambda f: f
    return fn


def vmap(fn, in_axes=0, out_axes=0):
    """Loop-and-stack vmap over the leading axis of every mapped argument (pytrees allowed);
    enough for scripts/train.py:559-579's per-sample predict."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = np.shape(tree_leaves(a)[0])[0]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else tree_map(lambda x: x[i], a) for a, ax in zip(args, axes)]
            outs.append(fn(*sl))
        return tree_map(lambda *xs: wrap(np.stack(xs, 0)), outs[0], *outs[1:])
    return mapped
