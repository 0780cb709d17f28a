"""Checkpoint ingest: parameter trees the reference writes -> the Flax-named dict
``HyperVLA.from_config(params=...)`` packs into device blobs.

Formats:
  * ``<dir>/<step>/EMA_params.pkl`` -- ``pickle.dump({"EMA_0.999": tree})`` (scripts/train.py:684-699); the eval
    loops load it in preference to the raw params (data/simpler/evaluate.py:441-443).  Leaves are jax Arrays or
    numpy arrays; jax is not needed to read them: jax pickles an Array as
    ``jax._src.array._reconstruct_array(np_reconstruct, args, state, aval_state)``, which is resolved here to the
    underlying numpy array by a restricted unpickler (nothing else from the stream is ever imported).
  * ``<dir>/params_<step>.npz`` -- flat "a/b/c" keys; written by ``HyperVLA.save_pretrained`` and by
    tools/convert_orbax_checkpoint.py, which must run where orbax/jax are installed (the reference's environment):
    the orbax PyTree checkpoint (``CheckpointManager.restore``, hypervla/model.py:208-214) is a tensorstore/OCDBT
    directory and is not parsed here.
"""
from __future__ import annotations

import io
import os
import pickle
from typing import Optional

import numpy as np

from . import metadata as M

_ALLOWED = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
    ("collections", "OrderedDict"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"),
}


def _reconstruct_jax_array(fun, args, arr_state, aval_state=None):
    """Stand-in for jax._src.array._reconstruct_array: rebuild the numpy value, skip the device_put."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


class _ParamUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        if name == "_reconstruct_array" and module.startswith("jax"):
            return _reconstruct_jax_array
        if module.startswith("flax.core") and name in ("FrozenDict", "freeze"):
            return dict
        raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a parameter pickle")


def load_ema_pickle(path: str, key: str = "EMA_0.999") -> dict:
    with open(path, "rb") as f:
        tree = _ParamUnpickler(io.BytesIO(f.read())).load()
    if isinstance(tree, dict) and key in tree:
        tree = tree[key]
    return _as_numpy_tree(tree)


def _as_numpy_tree(tree):
    if isinstance(tree, dict):
        return {str(k): _as_numpy_tree(v) for k, v in tree.items()}
    return np.asarray(tree)


def load_flat_npz(path: str) -> dict:
    flat = np.load(path)
    params: dict = {}
    for k in flat.files:
        M.set_path(params, tuple(k.split("/")), flat[k])
    return params


def latest_step(checkpoint_path: str) -> Optional[int]:
    steps = [int(n) for n in os.listdir(checkpoint_path) if n.isdigit() and os.path.isdir(os.path.join(checkpoint_path, n))]
    steps += [int(n[7:-4]) for n in os.listdir(checkpoint_path) if n.startswith("params_") and n.endswith(".npz") and n[7:-4].isdigit()]
    return max(steps) if steps else None


def load_params(checkpoint_path: str, step: Optional[int] = None) -> dict:
    """EMA pickle of the step if present, else the flat npz; ``step`` defaults to the latest (model.py:212)."""
    step = step if step is not None else latest_step(checkpoint_path)
    if step is not None:
        ema = os.path.join(checkpoint_path, str(step), "EMA_params.pkl")
        if os.path.exists(ema):
            return load_ema_pickle(ema)
        npz = os.path.join(checkpoint_path, f"params_{step}.npz")
        if os.path.exists(npz):
            return load_flat_npz(npz)
    cand = sorted(n for n in os.listdir(checkpoint_path) if n.startswith("params") and n.endswith(".npz"))
    if cand and step is None:
        return load_flat_npz(os.path.join(checkpoint_path, cand[-1]))
    raise FileNotFoundError(
        f"no <step>/EMA_params.pkl or params_<step>.npz under {checkpoint_path}; convert the orbax checkpoint with "
        "tools/convert_orbax_checkpoint.py in an environment that has orbax (see INTEGRATION.md)")
