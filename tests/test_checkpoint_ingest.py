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


# ---- orbax PyTree checkpoint, zarr-per-leaf layout -----------------------------------------------------------------------------
def _write_zarr_v2(path, arr, chunks, compressor=None, order="C", sep="."):
    """Independent writer following the zarr v2 spec: .zarray metadata + one file per chunk, every chunk FULL size (edge chunks
    padded with fill_value), laid out in C or F order, optionally compressed."""
    import gzip, json, zlib
    import pyarrow as pa
    os.makedirs(path)
    arr = np.asarray(arr)
    meta = {"zarr_format": 2, "shape": list(arr.shape), "chunks": list(chunks), "dtype": arr.dtype.str, "order": order,
            "fill_value": 0, "filters": None, "compressor": None if compressor is None else {"id": compressor, "level": 1}}
    if sep != ".":
        meta["dimension_separator"] = sep
    with open(os.path.join(path, ".zarray"), "w") as f:
        json.dump(meta, f)
    grid = [range(-(-s // c)) for s, c in zip(arr.shape, chunks)] if arr.ndim else [range(1)]
    import itertools
    for idx in itertools.product(*grid):
        block = np.zeros(chunks if arr.ndim else (), arr.dtype)
        if arr.ndim:
            sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, arr.shape))
            part = arr[sl]
            block[tuple(slice(0, n) for n in part.shape)] = part
        else:
            block[...] = arr
        raw = block.tobytes(order=order)
        if compressor == "zstd":
            raw = pa.Codec("zstd").compress(raw, asbytes=True)
        elif compressor == "gzip":
            raw = gzip.compress(raw)
        elif compressor == "zlib":
            raw = zlib.compress(raw)
        name = sep.join(str(i) for i in idx) if arr.ndim else "0"
        if sep == "/" and arr.ndim > 1:
            os.makedirs(os.path.join(path, os.path.dirname(name)), exist_ok=True)
        with open(os.path.join(path, name), "wb") as f:
            f.write(raw)


def test_orbax_zarr_per_leaf_checkpoint_is_read_without_orbax(tmp_path, params_p1):
    """hvla/orbax_reader.py: <ckpt>/<step>/default/<key.path>/ zarr v2 arrays -> the Flax-named pytree, for the chunkings /
    compressors / orders the format allows; load_params and latest_step pick it up."""
    from hvla import checkpoint as CK, metadata as M, orbax_reader as OR
    rng = np.random.default_rng(0)
    tree = {"task_pos_embedding": params_p1["task_pos_embedding"],                                  # (1,32,128) f32
            "context_encoder": {"encoderblock_0": {"MlpBlock_0": {"Dense_0": params_p1["context_encoder"]["encoderblock_0"]["MlpBlock_0"]["Dense_0"]}}},
            "odd": {"ragged": rng.standard_normal((5, 7)).astype(np.float32), "ints": np.arange(10, dtype=np.int32),
                    "scalar": np.float32(3.5), "half": rng.standard_normal((3, 4)).astype(np.float16)}}
    item = tmp_path / "ck" / "7000" / "default"
    variants = {"task_pos_embedding": ((1, 16, 128), "zstd", "C", "."), "kernel": ((128, 512), None, "C", "."), "bias": ((512,), "gzip", "C", "."),
                "ragged": ((2, 3), "zlib", "F", "."), "ints": ((4,), "zstd", "C", "."), "scalar": ((), None, "C", "."), "half": ((2, 4), None, "C", "/")}
    for path, v in M.iter_leaves(tree):
        ch, comp, order, sep = variants[path[-1]]
        _write_zarr_v2(str(item / ".".join(path)), v, ch, comp, order, sep)
    assert OR.is_orbax_step(str(tmp_path / "ck" / "7000"))
    got = CK.load_params(str(tmp_path / "ck"))                # latest step, raw params
    n = 0
    for path, v in M.iter_leaves(tree):
        g = M.get_path(got, path)
        assert g.dtype == np.asarray(v).dtype and g.shape == np.shape(v) and np.array_equal(g, v), path
        n += 1
    assert n == 7
    # an OCDBT checkpoint is recognised and refused with the way out, never mis-parsed
    oc = tmp_path / "ck2" / "100" / "default"
    os.makedirs(oc / "d")
    (oc / "manifest.ocdbt").write_bytes(b"\x0c\xdb\x3a\x2a")
    with pytest.raises(OR.OrbaxFormatError, match="convert_orbax_checkpoint"):
        CK.load_params(str(tmp_path / "ck2"), 100)
    # a chunk of the wrong size is an error, not silently accepted
    bad = tmp_path / "ck3" / "1" / "default"
    _write_zarr_v2(str(bad / "x"), np.arange(8, dtype=np.float32), (4,))
    with open(bad / "x" / "1", "wb") as f:
        f.write(b"\0" * 8)
    with pytest.raises(OR.OrbaxFormatError, match="bytes"):
        CK.load_params(str(tmp_path / "ck3"), 1)


def test_orbax_converter_script_against_a_stub_orbax(tmp_path, params_p1, monkeypatch):
    """tools/convert_orbax_checkpoint.py runs in the reference's environment; its logic (restore -> flatten with path -> flat npz
    that load_params reads back) is exercised here with stand-ins for the two library calls it makes."""
    import importlib.util
    from hvla import checkpoint as CK, metadata as M
    tree = {"task_token_projection": params_p1["task_token_projection"], "layer_pos_embedding": params_p1["layer_pos_embedding"]}

    class DictKey:
        def __init__(self, key):
            self.key = key

    def flatten_with_path(t, prefix=()):
        out = []
        for k in sorted(t):
            if isinstance(t[k], dict):
                out += flatten_with_path(t[k], prefix + (DictKey(k),))[0]
            else:
                out.append((prefix + (DictKey(k),), t[k]))
        return out, None

    class Manager:
        def __init__(self, path, handler):
            self.path = path

        def latest_step(self):
            return 4200

        def restore(self, step):
            assert step == 4200
            return tree
    jax = types.ModuleType("jax")
    jax.tree_util = types.SimpleNamespace(tree_flatten_with_path=flatten_with_path)
    orbax = types.ModuleType("orbax")
    orbax.checkpoint = types.ModuleType("orbax.checkpoint")
    orbax.checkpoint.CheckpointManager = Manager
    orbax.checkpoint.PyTreeCheckpointer = lambda: None
    for name, mod in (("jax", jax), ("orbax", orbax), ("orbax.checkpoint", orbax.checkpoint)):
        monkeypatch.setitem(sys.modules, name, mod)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("convert_orbax_checkpoint", os.path.join(root, "tools", "convert_orbax_checkpoint.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    monkeypatch.setattr(sys, "argv", ["convert_orbax_checkpoint.py", str(tmp_path)])
    conv.main()
    got = CK.load_params(str(tmp_path), 4200)
    for path, v in M.iter_leaves(tree):
        assert np.array_equal(M.get_path(got, path), v), path
