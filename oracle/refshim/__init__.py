"""refshim — run the reference's OWN hot-path Python (hypervla/model.py, hypervla/components/*.py) in this
container, where jax / flax / optax / orbax / tensorflow cannot be installed.

TEST INFRASTRUCTURE ONLY.  Nothing under hyper-vla_b200/ imports this; it is used by
tests/golden/make_ref_golden.py (to produce the committed reference-run fixtures tests/golden/ref_*.npz) and
by the CPU tests that re-run the reference when /root/reference is present.

`install()` registers lightweight stand-ins in sys.modules:
  jax, jax.numpy, jax.random, jax.tree_util, jax.lax, jax.typing   -> oracle/refshim/jaxlite.py (NumPy)
  flax, flax.linen, flax.struct, flax.core, flax.traverse_util      -> oracle/refshim/flaxlite.py (NumPy)
  transformers.FlaxDinov2Model                                       -> oracle/refshim/dinov2_torch.py (HF torch DINOv2)
  optax, distrax, orbax, tensorflow, ml_collections, ... and every other jax.* / flax.* submodule
                                                                     -> permissive stubs (unused on the path)
and puts the reference checkout on sys.path so `import hypervla.model` resolves to the reference's files.
What this pins and what it does not is stated in DESIGN.md §4: the HyperVLA-specific code is the reference's
own; Dense / LayerNorm / attention / gelu primitives are restated in flaxlite.py.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import types
import typing

import numpy as np

from . import flaxlite, jaxlite
from .stubs import Anything, StubFinder, StubModule

_INSTALLED = False
STUB_ROOTS = ("jax", "flax", "optax", "distrax", "orbax", "tensorflow", "tensorflow_probability", "tensorflow_hub",
              "tensorflow_datasets", "tensorflow_graphics", "ml_collections", "dlimp", "chex", "wandb", "absl",
              "transforms3d", "simpler_env")


def _module(name, **attrs):
    m = StubModule(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _struct_dataclass(cls=None, **kwargs):
    def wrap(c):
        c = dataclasses.dataclass(frozen=True)(c)
        c.replace = lambda self, **updates: dataclasses.replace(self, **updates)
        return c
    return wrap if cls is None else wrap(cls)


def _struct_field(pytree_node=True, **kwargs):
    return dataclasses.field(**kwargs)


def _flatten_dict(d, sep=None, prefix=()):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict) and v:
            out.update(_flatten_dict(v, sep, prefix + (k,)))
        else:
            out[prefix + (k,) if sep is None else sep.join(prefix + (k,))] = v
    return out


def _unflatten_dict(flat, sep=None):
    out = {}
    for k, v in flat.items():
        path = k.split(sep) if sep is not None else k
        t = out
        for p in path[:-1]:
            t = t.setdefault(p, {})
        t[path[-1]] = v
    return out


def _deep_dict(x):
    return {k: _deep_dict(v) for k, v in x.items()} if isinstance(x, dict) else x


def real_jax_available() -> bool:
    import importlib.util
    try:
        spec = importlib.util.find_spec("jax")
    except (ImportError, ValueError):
        return False
    return spec is not None and not isinstance(sys.modules.get("jax"), StubModule)


def install(reference_root: str = "/root/reference") -> None:
    """Idempotent.  Refuses to shadow a real jax (then the real reference should be run instead)."""
    global _INSTALLED
    if _INSTALLED:
        return
    if real_jax_available():
        raise RuntimeError("a real jax is importable: run the reference directly instead of through refshim")
    if not os.path.isdir(os.path.join(reference_root, "hypervla")):
        raise FileNotFoundError(f"reference checkout not found at {reference_root}")
    J, F = jaxlite, flaxlite

    jnp = J.make_jnp()
    sys.modules["jax.numpy"] = jnp
    random = _module("jax.random", PRNGKey=J.PRNGKey, key=J.PRNGKey, split=J.split, normal=J.random_normal,
                     uniform=J.random_uniform, KeyArray=np.ndarray)
    tree_util = _module("jax.tree_util", tree_map=J.tree_map, tree_map_with_path=J.tree_map_with_path,
                        tree_flatten=J.tree_flatten, tree_unflatten=J.tree_unflatten, tree_leaves=J.tree_leaves,
                        DictKey=J.DictKey, SequenceKey=J.SequenceKey)
    lax = _module("jax.lax", stop_gradient=lambda x: x, rsqrt=lambda x: J.wrap(1.0 / np.sqrt(np.asarray(x))))
    jtyping = _module("jax.typing", ArrayLike=typing.Any, DTypeLike=typing.Any)
    def _one_hot(x, num_classes, dtype=None):          # jax.nn.one_hot (BinTokenizer.decode, octo/model/components/tokenizers.py:272)
        x = np.asarray(x)
        return jaxlite.wrap((x[..., None] == np.arange(num_classes)).astype(dtype or np.float64)) if hasattr(jaxlite, "wrap") \
            else (x[..., None] == np.arange(num_classes)).astype(dtype or np.float64)
    jnn = _module("jax.nn", gelu=F.gelu, softmax=F.softmax, swish=F.swish, silu=F.swish, relu=F.relu, one_hot=_one_hot)
    mh = _module("jax.experimental.multihost_utils", process_allgather=lambda x, *a, **k: x)
    experimental = _module("jax.experimental", multihost_utils=mh)
    _module("jax", numpy=jnp, random=random, tree_util=tree_util, lax=lax, typing=jtyping, nn=jnn,
            experimental=experimental, Array=np.ndarray, jit=J.jit, vmap=J.vmap, tree_map=J.tree_map,
            tree_leaves=J.tree_leaves, tree_flatten=J.tree_flatten, tree_unflatten=J.tree_unflatten,
            device_get=lambda x: x, device_put=lambda x, *a, **k: x, process_index=lambda: 0, process_count=lambda: 1,
            __version__="0.4.20-refshim")

    linen = _module("flax.linen", Module=F.Module, compact=F.compact, nowrap=F.nowrap, Dense=F.Dense,
                    DenseGeneral=F.DenseGeneral, LayerNorm=F.LayerNorm, Dropout=F.Dropout,
                    MultiHeadDotProductAttention=F.MultiHeadDotProductAttention, SelfAttention=F.SelfAttention,
                    gelu=F.gelu, swish=F.swish, silu=F.silu, relu=F.relu, softmax=F.softmax, merge_param=F.merge_param,
                    initializers=F.initializers)
    sys.modules["flax.linen.initializers"] = F.initializers
    struct = _module("flax.struct", dataclass=_struct_dataclass, field=_struct_field)
    core = _module("flax.core", freeze=_deep_dict, unfreeze=_deep_dict, FrozenDict=dict)
    trav = _module("flax.traverse_util", flatten_dict=lambda d, sep=None, **k: _flatten_dict(d, sep),
                   unflatten_dict=lambda d, sep=None: _unflatten_dict(d, sep))
    _module("flax", linen=linen, struct=struct, core=core, traverse_util=trav, __version__="0.8.1-refshim")

    sys.meta_path.append(StubFinder(STUB_ROOTS))

    import transformers
    from . import dinov2_torch
    transformers.Dinov2Model, transformers.Dinov2Config     # resolve the lazy module first: it may re-register itself
    for mod in {id(m): m for m in (transformers, sys.modules["transformers"])}.values():
        mod.FlaxDinov2Model = dinov2_torch.FlaxDinov2Model
        mod.FlaxCLIPVisionModel = Anything("transformers.FlaxCLIPVisionModel")
        mod.FlaxT5EncoderModel = Anything("transformers.FlaxT5EncoderModel")
        mod.FlaxAutoModel = Anything("transformers.FlaxAutoModel")

    # einops probes every framework it finds in sys.modules (our stubs included): pin the numpy backend
    import einops._backends as eb
    eb._type2backend[jaxlite.Arr] = eb.NumpyBackend()
    eb._type2backend[np.ndarray] = eb._type2backend[jaxlite.Arr]

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    _INSTALLED = True


def reference_available(reference_root: str = "/root/reference") -> bool:
    return os.path.isdir(os.path.join(reference_root, "hypervla")) and not real_jax_available()
