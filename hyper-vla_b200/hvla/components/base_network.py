"""BaseNetwork: host-side mirror of ``hypervla/components/base_network.py`` (class ``BaseNetwork``
:11-183).  ``apply(..., method=BaseNetwork.predict_action)`` keeps the call shape of
``HyperVLA.sample_actions`` (model.py:125-136) and dispatches to ``hvla_act`` /
``hvla_act_host``: DINOv2 encoder -> per-task base ViT -> mix action head (base_vit.py:109-226,
action_heads.py:430-472, 524-538)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class BaseNetwork:
    model_type: str = "vit"
    action_head_type: str = "mix"
    octo_kwargs: dict = field(default_factory=dict)
    cnn_kwargs: dict = field(default_factory=dict)
    vit_kwargs: dict = field(default_factory=dict)
    action_head_kwargs: dict = field(default_factory=dict)
    action_horizon: int = 4
    action_dim: int = 7

    @classmethod
    def from_config(cls, config: dict) -> "BaseNetwork":
        return cls(**config["base_net_kwargs"], octo_kwargs=config.get("model", {}))

    def predict_action(self, variables, observation, task, timestep_pad_mask, rng=None, train=False,
                       image_embeddings=None, *, model=None, task_index=None, attention_maps=False):
        """base_network.py:170-183: squeeze the window axis, encode, head.  Returns (action, gripper_logits), and with
        ``attention_maps`` also the DINOv2 / base-encoder attention weights (the reference's sown intermediates)."""
        base_params = variables["params"]
        shape = tuple(observation.shape)
        if len(shape) == 5:
            if shape[1] != 1:
                raise ValueError("window_size must be 1 (the reference squeezes axis 1: model.py:117)")
            observation = observation.reshape(shape[0], *shape[2:]) if not hasattr(observation, "squeeze") else observation.squeeze(1)
        try:
            import torch
            on_device = torch.is_tensor(observation) and observation.is_cuda
        except ImportError:  # pragma: no cover
            on_device = False
        rt = model.runtime
        if self.action_head_type == "discrete":       # DiscreteActionHead.predict_action(argmax=True) + BinTokenizer.decode (action_heads.py:372-396)
            act, tok = rt.act_discrete(observation, base_params.weights, task_index)
            if on_device:
                return act, tok
            return act.cpu().numpy(), tok.cpu().numpy()
        if attention_maps:
            act, logit, dmaps, bmaps = rt.act_debug(observation, base_params.weights, task_index)
            if on_device:
                return act, logit, dmaps, bmaps
            return act.cpu().numpy(), logit.cpu().numpy(), dmaps.cpu().numpy(), bmaps.cpu().numpy()
        if on_device:
            return rt.act_device(observation, base_params.weights, task_index)
        return rt.act_host(observation, base_params.weights, task_index)

    def apply(self, variables, *args, method=None, model=None, mutable=None, rngs=None, **kwargs):
        if method is not None and getattr(method, "__name__", "") != "predict_action":
            raise ValueError("only BaseNetwork.predict_action is on the hot path")
        return self.predict_action(variables, *args, model=model, **kwargs)
