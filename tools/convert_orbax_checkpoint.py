#!/usr/bin/env python
"""Run this in the REFERENCE's environment (jax + flax + orbax installed), not in the hvla one:

    python convert_orbax_checkpoint.py <checkpoint_dir> [step]

Restores the orbax PyTree checkpoint the way HyperVLA.load_pretrained does (hypervla/model.py:208-214, without the
shape template) and writes <checkpoint_dir>/params_<step>.npz with flat "a/b/c" keys, which
hvla.model.HyperVLA.load_pretrained reads.  Needed only for OCDBT checkpoints (manifest.ocdbt): the zarr-per-leaf layout is read
directly by hvla/orbax_reader.py.  orbax is not installable in the build container; tests/test_checkpoint_ingest.py runs this
script against stand-ins for the two library calls it makes."""
import sys

import numpy as np


def main():
    import jax
    import orbax.checkpoint
    path = sys.argv[1]
    mgr = orbax.checkpoint.CheckpointManager(path, orbax.checkpoint.PyTreeCheckpointer())
    step = int(sys.argv[2]) if len(sys.argv) > 2 else mgr.latest_step()
    params = mgr.restore(step)
    flat = {}
    for keys, leaf in jax.tree_util.tree_flatten_with_path(params)[0]:
        flat["/".join(str(getattr(k, "key", k)) for k in keys)] = np.asarray(leaf)
    np.savez(f"{path}/params_{step}.npz", **flat)
    print(f"wrote {path}/params_{step}.npz ({len(flat)} leaves)")


if __name__ == "__main__":
    main()
