"""Synthetic inputs of the shapes the reference's callers produce.

There is no network in this environment (no datasets, no T5 / DINOv2 checkpoints), so the
language-token embeddings and initial-image embeddings that the eval loop computes upstream
(data/simpler/evaluate.py:240-277) are drawn from ``numpy.random.default_rng(1000 + config_index)``.
"""
from __future__ import annotations

import numpy as np

from . import config as C


def make_inputs(config_index: int, B: int, T: int, task_index=None) -> dict:
    rng = np.random.default_rng(1000 + config_index)
    tok = rng.standard_normal((T, C.LANG_TOKENS, C.LANG_DIM)).astype(np.float32)
    n_t = rng.integers(3, 21, size=T)
    am = (np.arange(C.LANG_TOKENS)[None, :] < n_t[:, None]).astype(np.int32)
    ids = (rng.integers(1, 32000, size=(T, C.LANG_TOKENS)) * am).astype(np.int32)
    patch = rng.standard_normal((T, C.DINO_TOKENS, C.DINO_DIM)).astype(np.float32)
    images = rng.integers(0, 256, size=(B, 1, C.IMAGE_SIZE, C.IMAGE_SIZE, 3), dtype=np.uint8)
    if task_index is None:
        if T == B:
            task_index = np.arange(B, dtype=np.int32)
        elif T == 1:
            task_index = np.zeros(B, dtype=np.int32)
        else:
            task_index = rng.permutation(np.repeat(np.arange(T), -(-B // T))[:B]).astype(np.int32)
    return {
        "instruction_dict": {"language_instruction": {"input_ids": ids, "attention_mask": am, "token_embedding": tok}},
        "initial_state": {"patch_embeddings": patch},
        "images": images,
        "task_index": np.asarray(task_index, np.int32),
        "timestep_pad_mask": np.ones((B, 1), dtype=np.float64),
    }
