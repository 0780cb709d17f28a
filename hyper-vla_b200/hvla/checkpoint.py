"""Checkpoint ingest: parameter trees the reference writes -> the Flax-named dict
``HyperVLA.from_config(params=...)`` packs into device blobs.

Formats:
  * ``<dir>/<step>/EMA_params.pkl`` -- ``pickle.dump({"EMA_0.999": tree})`` (scripts/train.py:684-699); the eval
    loops load it in preference to the raw params (data/simpler/evaluate.py:441-443).  Leaves are jax Arrays or
    numpy arrays; jax is not needed to read them: jax pickles an Array as
    ``jax._src.array._reconstruct_array(np_reconstruct, args, state, aval_state)``, which is resolved here to the
    underlying numpy array by a restricted unpickler (nothing else from the stream is ever imported).
  * ``<dir>/params_<step>.npz`` -- flat "a/b/c" keys; written by ``HyperVLA.save_pretrained`` and by
    tools/convert_orbax_checkpoint.py (runs where orbax/jax are installed: the reference's environment).
  * ``<dir>/<step>/default/<leaf.path>/.zarray`` -- the orbax PyTree checkpoint itself (``CheckpointManager.restore``,
    hypervla/model.py:208-214) in its zarr-per-leaf encoding, read without orbax by hvla/orbax_reader.py; the OCDBT encoding
    (``manifest.ocdbt``) is detected and refused with a pointer to the converter.
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


def ema_key(ema) -> str:
    """``--EMA 0.999`` of the eval loops -> the pickle key ``f"EMA_{args.EMA}"`` (data/simpler/evaluate.py:443)."""
    return ema if isinstance(ema, str) and ema.startswith("EMA_") else f"EMA_{ema}"


def load_ema_pickle(path: str, key: str = "EMA_0.999") -> dict:
    """One EMA parameter tree of ``pickle.dump({f"EMA_{c}": tree, ...})`` (scripts/train.py:684-699)."""
    with open(path, "rb") as f:
        trees = _ParamUnpickler(io.BytesIO(f.read())).load()
    if not isinstance(trees, dict) or key not in trees:
        have = sorted(str(k) for k in trees) if isinstance(trees, dict) else type(trees).__name__
        raise KeyError(f"{path} holds no {key} entry (found: {have}); pass ema=<coefficient> matching one of them")
    return _as_numpy_tree(trees[key])


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


def load_params(checkpoint_path: str, step: Optional[int] = None, ema=None) -> dict:
    """The RAW parameters of ``step`` (what ``HyperVLA.load_pretrained`` restores, model.py:209-214; ``step`` defaults to
    the latest, :212).  ``ema`` (e.g. ``0.999``) selects the EMA tree ``<step>/EMA_params.pkl[f"EMA_{ema}"]`` instead -- what
    the eval loops swap in only when run with ``--EMA`` (data/simpler/evaluate.py:439-444)."""
    step = step if step is not None else latest_step(checkpoint_path)
    if ema is not None:
        if step is None:
            raise FileNotFoundError(f"no step directory under {checkpoint_path} to read EMA_params.pkl from")
        pkl = os.path.join(checkpoint_path, str(step), "EMA_params.pkl")
        if not os.path.exists(pkl):
            raise FileNotFoundError(f"ema={ema!r} requested but {pkl} does not exist")
        return load_ema_pickle(pkl, ema_key(ema))
    if step is not None:
        npz = os.path.join(checkpoint_path, f"params_{step}.npz")
        if os.path.exists(npz):
            return load_flat_npz(npz)
        from . import orbax_reader as OR
        step_dir = os.path.join(checkpoint_path, str(step))
        if OR.is_orbax_step(step_dir):             # the orbax checkpoint itself (zarr-per-leaf layout; OCDBT is refused with a hint)
            return OR.read_orbax_pytree(step_dir)
    cand = sorted(n for n in os.listdir(checkpoint_path) if n.startswith("params") and n.endswith(".npz"))
    if cand and step is None:
        return load_flat_npz(os.path.join(checkpoint_path, cand[-1]))
    raise FileNotFoundError(
        f"no params_<step>.npz under {checkpoint_path}; convert the orbax checkpoint with tools/convert_orbax_checkpoint.py "
        "in an environment that has orbax (see INTEGRATION.md), or pass ema=<coefficient> to read <step>/EMA_params.pkl")
