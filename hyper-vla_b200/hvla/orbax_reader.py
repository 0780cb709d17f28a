"""Reads the parameters of an orbax PyTree checkpoint WITHOUT orbax / tensorstore / jax.

The reference saves with ``CheckpointManager(path, PyTreeCheckpointer()).save(step, params, save_args_from_target(params))``
(hypervla/model.py:232-256) and restores with ``checkpointer.restore(step, params_shape)`` (model.py:208-214).  On disk that is
``<path>/<step>/default/`` holding one array per leaf.  Two on-disk encodings exist, chosen by the orbax version that WROTE it:

* **one zarr (v2) directory per leaf** -- ``<item>/<key.path.joined.by.dots>/.zarray`` + chunk files (orbax < 0.5 default, or
  ``use_ocdbt=False``): parsed here from the public zarr v2 spec (``.zarray`` JSON: shape / chunks / dtype / order /
  compressor / fill_value / dimension_separator; a chunk file is the full-size chunk, C- or F-ordered, optionally compressed
  with zstd / gzip / zlib; missing chunks are fill_value).  ``save_args_from_target`` marks every leaf ``aggregate=False``, so no
  leaf hides in the msgpack aggregate file.
* **OCDBT** (``manifest.ocdbt`` + ``d/`` data files; the default of newer orbax): tensorstore's B+tree key-value store.  It is
  detected and refused with a pointer to tools/convert_orbax_checkpoint.py (run once where orbax exists); guessing at that
  binary format without a file to test against would be worse than saying so.

Pinned by tests/test_checkpoint_ingest.py against fixtures written by an independent zarr-v2 writer that follows the same spec.
"""
from __future__ import annotations

import itertools
import json
import os
import zlib
from typing import Dict

import numpy as np


class OrbaxFormatError(RuntimeError):
    pass


def _decompress(raw: bytes, compressor) -> bytes:
    if compressor is None:
        return raw
    cid = compressor.get("id")
    if cid == "zstd":
        try:
            import pyarrow as pa
        except ImportError as e:  # pragma: no cover
            raise OrbaxFormatError("zstd-compressed zarr chunk and no zstd codec available (pyarrow)") from e
        # zstd frames carry their content size; pyarrow wants it up front, so read it from the frame header
        return _zstd_decompress(raw, pa)
    if cid == "gzip":
        import gzip
        return gzip.decompress(raw)
    if cid == "zlib":
        return zlib.decompress(raw)
    raise OrbaxFormatError(f"zarr compressor {cid!r} is not supported here (zstd / gzip / zlib / none are); "
                           "convert the checkpoint with tools/convert_orbax_checkpoint.py")


def _zstd_frame_content_size(raw: bytes):
    """Frame_Content_Size of a zstd frame header (RFC 8878 section 3.1.1.1), or None when the frame does not record it."""
    if len(raw) < 6 or raw[:4] != b"\x28\xb5\x2f\xfd":
        raise OrbaxFormatError("not a zstd frame")
    fhd = raw[4]
    fcs_flag, single_segment, dict_flag = fhd >> 6, (fhd >> 5) & 1, fhd & 3
    pos = 5 + (0 if single_segment else 1) + (0, 1, 2, 4)[dict_flag]
    size = (1 if single_segment else 0, 2, 4, 8)[fcs_flag]
    if size == 0:
        return None
    val = int.from_bytes(raw[pos:pos + size], "little")
    return val + 256 if size == 2 else val


def _zstd_decompress(raw: bytes, pa) -> bytes:
    n = _zstd_frame_content_size(raw)
    codec = pa.Codec("zstd")
    if n is not None:
        return codec.decompress(raw, decompressed_size=n).to_pybytes()
    for guess in (1 << 20, 1 << 24, 1 << 28):      # frame without a recorded size: grow the output buffer
        try:
            return codec.decompress(raw, decompressed_size=guess).to_pybytes()
        except Exception:
            continue
    raise OrbaxFormatError("zstd chunk without a content size could not be decompressed")


def read_zarr_array(path: str) -> np.ndarray:
    """One zarr v2 array directory -> numpy array."""
    with open(os.path.join(path, ".zarray")) as f:
        meta = json.load(f)
    if meta.get("zarr_format") != 2:
        raise OrbaxFormatError(f"{path}: zarr_format {meta.get('zarr_format')!r} (only v2 is read here)")
    if meta.get("filters"):
        raise OrbaxFormatError(f"{path}: zarr filters are not supported")
    shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
    dtype = np.dtype(meta["dtype"])
    order = meta.get("order", "C")
    sep = meta.get("dimension_separator", ".")
    fill = meta.get("fill_value")
    out = np.empty(shape, dtype)
    if fill is not None:
        out[...] = np.array(fill if not isinstance(fill, str) else float(fill), dtype)
    else:
        out[...] = 0
    if len(shape) == 0:                                # scalar: a single chunk named "0"
        p = os.path.join(path, "0")
        if os.path.exists(p):
            with open(p, "rb") as f:
                out[...] = np.frombuffer(_decompress(f.read(), meta.get("compressor")), dtype, count=1)[0]
        return out
    grid = [range(-(-s // c)) for s, c in zip(shape, chunks)]
    for idx in itertools.product(*grid):
        p = os.path.join(path, sep.join(str(i) for i in idx))
        if not os.path.exists(p):
            continue                                   # missing chunk = fill_value
        with open(p, "rb") as f:
            raw = _decompress(f.read(), meta.get("compressor"))
        if len(raw) != int(np.prod(chunks)) * dtype.itemsize:
            raise OrbaxFormatError(f"{p}: chunk has {len(raw)} bytes, expected {int(np.prod(chunks)) * dtype.itemsize}")
        block = np.frombuffer(raw, dtype).reshape(chunks, order=order)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out


def item_directory(step_dir: str) -> str:
    """``<step>/default`` (CheckpointManager with one un-named item) or the step directory itself."""
    d = os.path.join(step_dir, "default")
    return d if os.path.isdir(d) else step_dir


def is_orbax_step(step_dir: str) -> bool:
    d = item_directory(step_dir)
    if not os.path.isdir(d):
        return False
    if os.path.exists(os.path.join(d, "manifest.ocdbt")):
        return True
    return any(os.path.exists(os.path.join(d, n, ".zarray")) for n in os.listdir(d))


def read_orbax_pytree(step_dir: str) -> Dict:
    """Parameters of ``<checkpoint>/<step>`` as a nested dict of numpy arrays (Flax names)."""
    d = item_directory(step_dir)
    if os.path.exists(os.path.join(d, "manifest.ocdbt")):
        raise OrbaxFormatError(
            f"{d} is an OCDBT checkpoint (manifest.ocdbt): tensorstore's key-value store is not parsed here.  Run "
            "tools/convert_orbax_checkpoint.py once in the reference's environment (it writes params_<step>.npz), or re-save with "
            "PyTreeCheckpointHandler(use_ocdbt=False)")
    leaves = sorted(n for n in os.listdir(d) if os.path.exists(os.path.join(d, n, ".zarray")))
    if not leaves:
        raise OrbaxFormatError(f"no zarr arrays under {d}")
    tree: Dict = {}
    for name in leaves:
        node = tree
        keys = name.split(".")
        for k in keys[:-1]:
            node = node.setdefault(k, {})
            if not isinstance(node, dict):
                raise OrbaxFormatError(f"{name}: key path collides with a leaf")
        node[keys[-1]] = read_zarr_array(os.path.join(d, name))
    return tree
