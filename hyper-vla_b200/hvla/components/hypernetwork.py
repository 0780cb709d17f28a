"""HyperNetwork: host-side mirror of ``hypervla/components/hypernetwork.py`` (class
``HyperNetwork`` :28-242).  ``apply`` keeps the Flax call shape used by ``HyperVLA.create_tasks``
(model.py:73-80) and dispatches to ``hvla_generate`` (context encoder + all 73 output heads)."""
from __future__ import annotations

from enum import IntEnum
from typing import Dict

import numpy as np


class InitOptions(IntEnum):          # hypernetwork.py:24-26
    BIAS_INIT = 0
    VARIANCE_INIT = 1


class HyperNetwork:
    def __init__(self, base_net_metadata: Dict, hypernet_kwargs: Dict):
        self.base_net_metadata = base_net_metadata
        self.hypernet_kwargs = hypernet_kwargs
        self.layer_token_num = base_net_metadata["block_num"]      # generation_strategy == 'block'

    def apply(self, variables, tasks, train: bool = False, initial_states=None, *, model=None, task_ids=None, base_params=None,
              **_unused):
        """-> ((base_params, context_embedding), intermediate_states)   (hypernetwork.py:199-219)"""
        if train:
            raise ValueError("hvla is an inference path: train=True is unsupported")
        if model is None:
            raise ValueError("HyperNetwork.apply needs the owning HyperVLA (model=...) for its device runtime")
        if variables["params"] is not model.params:
            raise ValueError("HyperNetwork.apply: variables['params'] must be the owning model's params")
        if initial_states is None or "patch_embeddings" not in initial_states:
            raise ValueError("use_initial_image=True: initial_state['patch_embeddings'] is required")   # hypernetwork.py:118-126
        from ..model import GeneratedBaseParams
        lang = tasks["language_instruction"]
        pe = initial_states["patch_embeddings"]
        cls = pe[:, 0]                                                   # initial_image[:, :1]  (:126)
        pad = tasks.get("pad_mask_dict", {}).get("language_instruction")
        if task_ids is not None:      # task-switch scheduler: regenerate only these rows of the existing buffers, in place
            model.runtime.generate(lang["token_embedding"], lang["attention_mask"], cls, pad, rows=task_ids,
                                   into=(base_params.weights, base_params.context_embedding))
            base_params._tree = None
            base_params.generation += 1
            ctx = base_params.context_embedding
            return (base_params, ctx.reshape(int(ctx.shape[0]), 1, -1)), {}
        weights, ctx = model.runtime.generate(lang["token_embedding"], lang["attention_mask"], cls, pad)
        T = int(weights.shape[0])
        base_params = GeneratedBaseParams(model, weights, ctx, squeeze=(T == 1))
        return (base_params, ctx.reshape(T, 1, -1)), {}
