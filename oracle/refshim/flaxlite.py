"""A small NumPy stand-in for the slice of `flax.linen` (0.8.1) the reference's hot-path modules use.

TEST INFRASTRUCTURE (see oracle/refshim/__init__.py).  It lets the reference's own, unmodified
`hypervla/components/*.py` and `hypervla/model.py` be imported and *executed* in this container, where
jax/flax are not installable.  What is restated here (from flax 0.8.1's documented behaviour, sources not on
this box) are the framework primitives only:

* the Module system: dataclass-style fields, lazy `setup`, `@compact`, scope paths and auto-names
  (`Dense_0`, attribute names for `setup` children, `attr_key` for dict-valued attributes),
  `param`, `sow`, `apply(..., method=, mutable=)`, `init`;
* `Dense`, `DenseGeneral`, `LayerNorm` (eps 1e-6, fast variance), `MultiHeadDotProductAttention`
  (q/sqrt(d) before QK^T, `where(mask, s, finfo.min)`, softmax), `Dropout` (identity when deterministic),
  `gelu` (tanh approximation), `swish`.
Everything HyperVLA-specific (token layout, masks, heads, reshapes, action squashing) is the reference's code.
"""
from __future__ import annotations

import inspect
import typing

import numpy as np

from . import jaxlite as J
from .stubs import Anything

_STACK = []                       # [(module, mode)]  mode in {"setup", "compact", "method"}


class _Run:
    def __init__(self, variables, mutable, init, rngs):
        self.variables = variables
        self.mutable = mutable
        self.init = init
        self.rngs = rngs or {}
        self.state = {}
        self.rng_counter = 0

    def is_mutable(self, col):
        if self.mutable is True:
            return True
        if not self.mutable:
            return False
        return col in self.mutable


def _get_in(tree, path):
    for p in path:
        if not isinstance(tree, dict) or p not in tree:
            return None
        tree = tree[p]
    return tree


def _set_in(tree, path, value):
    for p in path[:-1]:
        tree = tree.setdefault(p, {})
    tree[path[-1]] = value


def compact(fn):
    fn._flaxlite_compact = True
    return fn


def nowrap(fn):
    fn._flaxlite_nowrap = True
    return fn


def _wrap_method(fn):
    is_compact = getattr(fn, "_flaxlite_compact", False)

    def method(self, *args, **kwargs):
        if not isinstance(self, Module):
            return fn(self, *args, **kwargs)
        self._try_setup()
        if is_compact:
            self._autonames = {}
        _STACK.append((self, "compact" if is_compact else "method"))
        try:
            return fn(self, *args, **kwargs)
        finally:
            _STACK.pop()
    method.__name__ = fn.__name__
    method.__qualname__ = getattr(fn, "__qualname__", fn.__name__)
    method.__doc__ = fn.__doc__
    method.__wrapped__ = fn
    return method


class Module:
    """flax.linen.Module stand-in (see module docstring)."""
    _fields: tuple = ()

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        fields = []
        for klass in reversed(cls.__mro__):
            if klass in (object, Module):
                continue
            for name, ann in vars(klass).get("__annotations__", {}).items():
                if name in ("name", "parent") or name.startswith("_"):
                    continue
                if typing.get_origin(ann) is typing.ClassVar or (isinstance(ann, str) and ann.startswith("ClassVar")):
                    continue
                if name not in fields:
                    fields.append(name)
        cls._fields = tuple(fields)
        for name, obj in list(vars(cls).items()):
            if name in fields or name == "setup" or not inspect.isfunction(obj):
                continue
            if name.startswith("_") and name != "__call__":
                continue
            if getattr(obj, "_flaxlite_nowrap", False):
                continue
            setattr(cls, name, _wrap_method(obj))

    def __init__(self, *args, **kwargs):
        d = object.__getattribute__(self, "__dict__")
        d["_setup_done"] = False
        d["_in_setup"] = False
        d["_autonames"] = {}
        d["_run"] = None
        d["_path"] = None
        d["name"] = kwargs.pop("name", None)
        d["parent"] = kwargs.pop("parent", None)
        names = list(type(self)._fields)
        if len(args) > len(names):
            raise TypeError(f"{type(self).__name__}: too many positional arguments")
        given = dict(zip(names, args))
        for k, v in kwargs.items():
            if k not in names:
                raise TypeError(f"{type(self).__name__}: unexpected field {k!r}")
            if k in given:
                raise TypeError(f"{type(self).__name__}: field {k!r} given twice")
            given[k] = v
        for n in names:
            if n in given:
                d[n] = given[n]
            elif hasattr(type(self), n):
                d[n] = getattr(type(self), n)
            else:
                raise TypeError(f"{type(self).__name__}: missing field {n!r}")
        if _STACK:
            parent, mode = _STACK[-1]
            if mode == "compact":
                parent._adopt(self, d["name"] or parent._auto_name(type(self).__name__))
            elif mode == "setup" and d["name"] is not None:
                parent._adopt(self, d["name"])
            elif mode == "setup":
                d["parent"] = parent          # named by the attribute it is assigned to

    # ---- binding -------------------------------------------------------------------------
    def _auto_name(self, prefix):
        i = self._autonames.get(prefix, 0)
        self._autonames[prefix] = i + 1
        return f"{prefix}_{i}"

    def _adopt(self, child, name):
        d = child.__dict__
        d["name"], d["parent"] = name, self
        d["_run"], d["_path"] = self._run, self._path + (name,)

    def _bound(self):
        return self._run is not None

    def _clone(self):
        saved = list(_STACK)
        del _STACK[:]
        try:
            return type(self)(**{n: self.__dict__[n] for n in type(self)._fields})
        finally:
            _STACK.extend(saved)

    def _try_setup(self):
        d = self.__dict__
        if d["_setup_done"] or d["_in_setup"] or not self._bound():
            return
        d["_in_setup"] = True
        _STACK.append((self, "setup"))
        try:
            self.setup()
        finally:
            _STACK.pop()
            d["_in_setup"] = False
        d["_setup_done"] = True

    def setup(self):
        pass

    def __getattr__(self, name):
        # only reached when normal lookup fails: attributes defined in setup()
        d = object.__getattribute__(self, "__dict__")
        if name.startswith("__") or d.get("_setup_done") or d.get("_in_setup") or d.get("_run") is None:
            raise AttributeError(f"{type(self).__name__!s} has no attribute {name!r}")
        self._try_setup()
        if name in d:
            return d[name]
        raise AttributeError(f"{type(self).__name__!s} has no attribute {name!r}")

    def __setattr__(self, name, value):
        d = self.__dict__
        if d.get("_in_setup"):
            self._name_children(name, value)
        d[name] = value

    def _name_children(self, attr, value):
        if isinstance(value, Module):
            if value.__dict__.get("_path") is None:
                self._adopt(value, value.name or attr)
        elif isinstance(value, dict):
            for k, v in value.items():
                self._name_children(f"{attr}_{k}", v)
        elif isinstance(value, (list, tuple)):
            for i, v in enumerate(value):
                self._name_children(f"{attr}_{i}", v)

    # ---- variables -----------------------------------------------------------------------
    def param(self, name, init_fn, *init_args, **init_kwargs):
        run = self._run
        if run is None:
            raise RuntimeError("param() on an unbound module")
        path = ("params",) + self._path + (name,)
        value = _get_in(run.variables, path)
        if value is None:
            if not run.init:
                raise KeyError("missing parameter " + "/".join(path[1:]))
            value = np.asarray(init_fn(self.make_rng("params"), *init_args, **init_kwargs))
            _set_in(run.variables, path, value)
        return J.wrap(np.asarray(value))

    def params_subtree(self):
        return _get_in(self._run.variables, ("params",) + self._path)

    def set_params_subtree(self, tree):
        _set_in(self._run.variables, ("params",) + self._path, tree)

    def has_variable(self, col, name):
        return _get_in(self._run.variables, (col,) + self._path + (name,)) is not None

    def sow(self, col, name, value, **_kwargs):
        run = self._run
        if run is None or not run.is_mutable(col):
            return False
        path = (col,) + self._path + (name,)
        prev = _get_in(run.state, path) or ()
        _set_in(run.state, path, tuple(prev) + (value,))
        return True

    def make_rng(self, name="params"):
        run = self._run
        run.rng_counter += 1
        base = run.rngs.get(name) if isinstance(run.rngs, dict) else run.rngs
        if base is None:
            base = J.PRNGKey(0)
        return J.split(base, run.rng_counter + 1)[-1]

    def is_initializing(self):
        return bool(self._run and self._run.init)

    def is_mutable_collection(self, col):
        return bool(self._run and self._run.is_mutable(col))

    # ---- entry points --------------------------------------------------------------------
    def _run_with(self, run, args, kwargs, method):
        top = self._clone()
        top.__dict__["_run"], top.__dict__["_path"] = run, ()
        if method is None:
            fn = type(top).__call__
        elif isinstance(method, str):
            fn = getattr(type(top), method)
        else:
            fn = method
        saved = list(_STACK)
        del _STACK[:]
        try:
            return fn(top, *args, **kwargs)
        finally:
            del _STACK[:]
            _STACK.extend(saved)

    def apply(self, variables, *args, rngs=None, method=None, mutable=False, capture_intermediates=False, **kwargs):
        run = _Run(dict(variables), mutable, False, rngs)
        out = self._run_with(run, args, kwargs, method)
        if mutable:
            return out, run.state
        return out

    def init(self, rngs, *args, method=None, mutable=True, **kwargs):
        if not isinstance(rngs, dict):
            rngs = {"params": rngs}
        run = _Run({"params": {}}, True, True, rngs)
        self._run_with(run, args, kwargs, method)
        return run.variables

    def init_with_output(self, rngs, *args, method=None, **kwargs):
        if not isinstance(rngs, dict):
            rngs = {"params": rngs}
        run = _Run({"params": {}}, True, True, rngs)
        out = self._run_with(run, args, kwargs, method)
        return out, run.variables

    def tabulate(self, *args, **kwargs):
        return f"<{type(self).__name__}: tabulate() not available in the shim>"

    def __repr__(self):
        return f"{type(self).__name__}(name={self.__dict__.get('name')!r})"


# ----------------------------------------------------------------------------------------------
# initializers (numpy draws; jax's random stream is not reproduced — values are random either way)
# ----------------------------------------------------------------------------------------------
class _Initializers:
    @staticmethod
    def zeros(key, shape, dtype=np.float64):
        return np.zeros(shape, dtype)

    @staticmethod
    def ones(key, shape, dtype=np.float64):
        return np.ones(shape, dtype)

    zeros_init = staticmethod(lambda: _Initializers.zeros)
    ones_init = staticmethod(lambda: _Initializers.ones)

    @staticmethod
    def constant(value):
        return lambda key, shape, dtype=np.float64: np.full(shape, value, dtype)

    @staticmethod
    def normal(stddev=1e-2):
        return lambda key, shape, dtype=np.float64: J.rng_of(key).standard_normal(tuple(shape)) * stddev

    @staticmethod
    def truncated_normal(stddev=1e-2, lower=-2.0, upper=2.0):
        def init(key, shape, dtype=np.float64):
            return np.clip(J.rng_of(key).standard_normal(tuple(shape)), lower, upper) * stddev
        return init

    @staticmethod
    def variance_scaling(scale, mode, distribution, in_axis=-2, out_axis=-1, **_kw):
        def init(key, shape, dtype=np.float64):
            shape = tuple(shape)
            rf = int(np.prod(shape)) // max(1, shape[in_axis] * shape[out_axis]) if len(shape) > 1 else 1
            fan_in = shape[in_axis] * rf if len(shape) > 1 else shape[0]
            fan_out = shape[out_axis] * rf if len(shape) > 1 else shape[0]
            denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
            var = scale / max(1.0, denom)
            rng = J.rng_of(key)
            if distribution == "uniform":
                lim = np.sqrt(3 * var)
                return rng.uniform(-lim, lim, shape)
            return rng.standard_normal(shape) * np.sqrt(var)
        return init

    @staticmethod
    def xavier_uniform(**kw):
        return _Initializers.variance_scaling(1.0, "fan_avg", "uniform", **kw)

    glorot_uniform = xavier_uniform

    @staticmethod
    def xavier_normal(**kw):
        return _Initializers.variance_scaling(1.0, "fan_avg", "normal", **kw)

    @staticmethod
    def lecun_normal(**kw):
        return _Initializers.variance_scaling(1.0, "fan_in", "normal", **kw)

    @staticmethod
    def he_normal(**kw):
        return _Initializers.variance_scaling(2.0, "fan_in", "normal", **kw)

    kaiming_normal = he_normal

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Anything(f"initializers.{name}")


initializers = _Initializers()


# ----------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------
def gelu(x, approximate=True):
    """flax.linen.gelu: default is the tanh approximation."""
    x = np.asarray(x)
    if approximate:
        cdf = 0.5 * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * (x ** 3))))
        return J.wrap(x * cdf)
    from math import erf
    return J.wrap(x * 0.5 * (1.0 + np.vectorize(erf)(x / np.sqrt(2.0))))


def swish(x):
    x = np.asarray(x)
    return J.wrap(x / (1.0 + np.exp(-x)))


silu = swish


def relu(x):
    return J.wrap(np.maximum(np.asarray(x), 0))


def softmax(x, axis=-1):
    x = np.asarray(x)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return J.wrap(e / e.sum(axis=axis, keepdims=True))


def merge_param(name, a, b):
    if a is None and b is None:
        raise ValueError(f"no value for {name}")
    if a is not None and b is not None:
        raise ValueError(f"{name} given twice")
    return b if a is None else a


class Dense(Module):
    features: int
    use_bias: bool = True
    dtype: typing.Any = None
    param_dtype: typing.Any = np.float64
    precision: typing.Any = None
    kernel_init: typing.Any = initializers.lecun_normal()
    bias_init: typing.Any = initializers.zeros

    @compact
    def __call__(self, inputs):
        inputs = np.asarray(inputs)
        kernel = self.param("kernel", self.kernel_init, (inputs.shape[-1], self.features))
        y = inputs @ np.asarray(kernel)
        if self.use_bias:
            y = y + np.asarray(self.param("bias", self.bias_init, (self.features,)))
        return J.wrap(y)


class DenseGeneral(Module):
    features: typing.Any
    axis: typing.Any = -1
    batch_dims: typing.Any = ()
    use_bias: bool = True
    dtype: typing.Any = None
    param_dtype: typing.Any = np.float64
    kernel_init: typing.Any = initializers.lecun_normal()
    bias_init: typing.Any = initializers.zeros
    precision: typing.Any = None

    @compact
    def __call__(self, inputs):
        inputs = np.asarray(inputs)
        feats = tuple(self.features) if isinstance(self.features, (tuple, list)) else (self.features,)
        axes = tuple(self.axis) if isinstance(self.axis, (tuple, list)) else (self.axis,)
        axes = tuple(a % inputs.ndim for a in axes)
        assert axes == tuple(range(inputs.ndim - len(axes), inputs.ndim)), "shim supports trailing contraction axes only"
        in_shape = tuple(inputs.shape[a] for a in axes)

        def kernel_init(key, shape):
            flat = (int(np.prod(in_shape)), int(np.prod(feats)))
            return np.asarray(self.kernel_init(key, flat)).reshape(shape)
        kernel = np.asarray(self.param("kernel", kernel_init, in_shape + feats))
        y = np.tensordot(inputs, kernel, axes=(axes, tuple(range(len(axes)))))
        if self.use_bias:
            y = y + np.asarray(self.param("bias", self.bias_init, feats))
        return J.wrap(y)


class LayerNorm(Module):
    epsilon: float = 1e-6
    dtype: typing.Any = None
    param_dtype: typing.Any = np.float64
    use_bias: bool = True
    use_scale: bool = True
    bias_init: typing.Any = initializers.zeros
    scale_init: typing.Any = initializers.ones
    reduction_axes: typing.Any = -1
    feature_axes: typing.Any = -1
    use_fast_variance: bool = True

    @compact
    def __call__(self, x):
        x = np.asarray(x)
        mean = x.mean(axis=-1, keepdims=True)
        if self.use_fast_variance:
            var = np.maximum(0.0, (x * x).mean(axis=-1, keepdims=True) - mean * mean)
        else:
            var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
        y = x - mean
        mul = 1.0 / np.sqrt(var + self.epsilon)
        if self.use_scale:
            mul = mul * np.asarray(self.param("scale", self.scale_init, (x.shape[-1],)))
        y = y * mul
        if self.use_bias:
            y = y + np.asarray(self.param("bias", self.bias_init, (x.shape[-1],)))
        return J.wrap(y)


class Dropout(Module):
    rate: float = 0.0
    broadcast_dims: typing.Any = ()
    deterministic: typing.Any = None
    rng_collection: str = "dropout"

    @compact
    def __call__(self, inputs, deterministic=None, rng=None):
        det = merge_param("deterministic", self.deterministic, deterministic)
        if det or self.rate == 0.0 or self.is_initializing():
            return inputs
        raise NotImplementedError("flaxlite: stochastic dropout is outside the inference path")


class MultiHeadDotProductAttention(Module):
    num_heads: int
    dtype: typing.Any = None
    param_dtype: typing.Any = np.float64
    qkv_features: typing.Any = None
    out_features: typing.Any = None
    broadcast_dropout: bool = True
    dropout_rate: float = 0.0
    deterministic: typing.Any = None
    precision: typing.Any = None
    kernel_init: typing.Any = initializers.lecun_normal()
    bias_init: typing.Any = initializers.zeros
    use_bias: bool = True
    decode: bool = False
    normalize_qk: bool = False

    @compact
    def __call__(self, inputs_q, inputs_k=None, inputs_v=None, *, inputs_kv=None, mask=None,
                 deterministic=None, dropout_rng=None, sow_weights=False):
        if inputs_kv is not None:
            inputs_k = inputs_v = inputs_kv
        if inputs_k is None:
            inputs_k = inputs_q
        if inputs_v is None:
            inputs_v = inputs_k
        inputs_q = np.asarray(inputs_q)
        features = self.out_features or inputs_q.shape[-1]
        qkv = self.qkv_features or inputs_q.shape[-1]
        assert qkv % self.num_heads == 0
        head_dim = qkv // self.num_heads

        def dense(name):
            return DenseGeneral(features=(self.num_heads, head_dim), axis=-1, kernel_init=self.kernel_init,
                                bias_init=self.bias_init, use_bias=self.use_bias, name=name)
        query = np.asarray(dense("query")(inputs_q))
        key = np.asarray(dense("key")(inputs_k))
        value = np.asarray(dense("value")(inputs_v))
        if self.dropout_rate > 0.0:
            det = merge_param("deterministic", self.deterministic, deterministic)
            if not det and not self.is_initializing():
                raise NotImplementedError("flaxlite: attention dropout is outside the inference path")
        # flax.linen.attention.dot_product_attention_weights
        query = query / np.sqrt(head_dim).astype(query.dtype)
        weights = np.einsum("...qhd,...khd->...hqk", query, key)
        if mask is not None:
            big_neg = np.finfo(weights.dtype).min
            weights = np.where(np.asarray(mask).astype(bool), weights, big_neg)
        weights = np.asarray(softmax(weights, axis=-1))
        if sow_weights:
            self.sow("intermediates", "attention_weights", J.wrap(weights))
        x = np.einsum("...hqk,...khd->...qhd", weights, value)
        out = DenseGeneral(features=features, axis=(-2, -1), kernel_init=self.kernel_init, bias_init=self.bias_init,
                           use_bias=self.use_bias, name="out")(x)
        return out


SelfAttention = MultiHeadDotProductAttention
