"""HyperVLA model API on the B200-native hot path.

Mirrors ``hypervla/model.py`` of the reference (class ``HyperVLA``: ``create_tasks`` :35-83,
``sample_actions`` :85-137, ``from_config`` :286-368, ``load_pretrained`` :139-224) -- same
names, argument order, return arity and parameter pytree -- but generate and act run in
libhvla.so (hand-written sm_100a CUDA) instead of Flax ``apply``.

Differences a caller can see (all documented in INTEGRATION.md):
  * ``base_params`` returned by ``create_tasks`` is a ``GeneratedBaseParams``: a read-only
    mapping that materialises the Flax-named pytree lazily, and carries the packed device
    blob the act kernels consume.  Callers of the reference only store it and hand it back
    (data/utils/hypervla_interface.py:144-146, 202), which keeps working.
  * batching: the reference is batch-1 at inference (hypervla_interface.py:207, 249).  Here
    ``create_tasks`` accepts T tasks and ``sample_actions`` B images plus an optional
    ``task_index`` (B,) -- semantics of the vmapped validation path (scripts/train.py:546-583).
  * ``intermediate_states`` holds ``{"gripper_logits": (B,4)}``; the sown attention maps (38 MB per image) are produced only on
    request: ``sample_actions(..., return_attention_maps=True)``.
"""
from __future__ import annotations

import json
import os
from collections.abc import Mapping
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from . import config as Cfg
from . import metadata as M
from . import params as P
from .components.base_network import BaseNetwork
from .components.hypernetwork import HyperNetwork


class GeneratedBaseParams(Mapping):
    """Base-net parameters produced by the hypernetwork for T tasks."""

    def __init__(self, model: "HyperVLA", weights, ctx_emb, squeeze: bool):
        self._model = model
        self.weights = weights          # torch tensor [T, NGP] on the model's device (packed rows)
        self.context_embedding = ctx_emb
        self.num_tasks = int(weights.shape[0])
        self._squeeze = squeeze
        self._tree = None
        self.task_index = None          # optional default env -> task map for sample_actions
        self.generation = 0             # bumped by every in-place row regeneration (task-switch scheduler)

    def packed_numpy(self) -> np.ndarray:
        return self.weights.float().cpu().numpy()[:, :M.n_generated(self._model.head_spec)]

    def tree(self) -> dict:
        """The pytree ``create_tasks`` returns in the reference: generated leaves (leading T unless
        T == 1, model.py:81) plus the shared DINOv2 leaves under encoder/image_encoder."""
        if self._tree is None:
            rows = self.packed_numpy()
            gen = P.unpack_generated(rows[0] if self._squeeze else rows, self._model.head_spec)
            gen.setdefault("encoder", {})["image_encoder"] = P.dino_tree_from_params(self._model.params)
            self._tree = gen
        return self._tree

    def __getitem__(self, k):
        return self.tree()[k]

    def __iter__(self):
        return iter(self.tree())

    def __len__(self):
        return len(self.tree())


def _seed_from_rng(rng, default: int) -> int:
    if rng is None:
        return default
    arr = np.asarray(rng).ravel()
    return int(arr[-1]) if arr.size else default


@dataclass
class HyperVLA:
    hypernet: HyperNetwork
    base_net: BaseNetwork
    config: dict
    params: dict
    base_net_metadata: dict
    example_batch: Optional[dict] = None
    dataset_statistics: Optional[dict] = None
    precision: str = "bf16"
    device: Any = None
    _runtime: Any = field(default=None, repr=False)

    # ---- construction ------------------------------------------------------------------------------
    @classmethod
    def from_config(cls, config: dict, example_batch: Optional[dict] = None, rng=None,
                    dataset_statistics: Optional[dict] = None, *, precision: str = "bf16", device=None,
                    params_variant: str = "P0", params: Optional[dict] = None) -> "HyperVLA":
        """Fresh (synthetic) weights from a config (reference: model.py:286-368).  ``params_variant``
        "P0" reproduces the reference's BIAS_INIT state; "P1" is the trained-like set used for parity."""
        cfg = Cfg.validate_config(config)
        meta = M.build_base_net_metadata(cfg)
        if params is None:
            params = P.init_params(_seed_from_rng(rng, int(cfg.get("seed", 2025))), params_variant, M.HeadSpec.from_config(cfg))
        return cls(hypernet=HyperNetwork(meta, cfg["hypernet_kwargs"]), base_net=BaseNetwork.from_config(cfg),
                   config=cfg, params=params, base_net_metadata=meta, example_batch=example_batch,
                   dataset_statistics=dataset_statistics, precision=precision, device=device)

    @classmethod
    def load_pretrained(cls, checkpoint_path: str, step: Optional[int] = None, *, precision: str = "bf16",
                        device=None, ema=None) -> "HyperVLA":
        """Reads ``config.json`` / ``dataset_statistics.json`` like the reference (model.py:152-189).
        Parameters (hvla/checkpoint.py): the RAW params of ``step`` as the reference restores them (model.py:209-214), from a
        flat ``params_<step>.npz`` ("a/b/c" keys, Flax names) written by ``save_pretrained`` / tools/convert_orbax_checkpoint.py;
        ``ema=0.999`` selects ``<step>/EMA_params.pkl["EMA_0.999"]`` instead -- the swap the eval loops do under ``--EMA``
        (data/simpler/evaluate.py:439-444, scripts/train.py:697-699), read without jax."""
        with open(os.path.join(checkpoint_path, "config.json")) as f:
            config = json.load(f)
        stats = None
        sp = os.path.join(checkpoint_path, "dataset_statistics.json")
        if os.path.exists(sp):
            with open(sp) as f:
                stats = json.load(f)
        from . import checkpoint as CK
        params = CK.load_params(checkpoint_path, step, ema=ema)
        if "shared_modules" in config.get("hypernet_kwargs", {}):
            config["hypernet_kwargs"]["shared_modules"] = tuple(config["hypernet_kwargs"]["shared_modules"])
        return cls.from_config(config, None, None, stats, precision=precision, device=device, params=params)

    def save_pretrained(self, step: int, checkpoint_path: str) -> None:
        os.makedirs(checkpoint_path, exist_ok=True)
        flat = {}
        for path, v in M.iter_leaves(self.params):
            flat["/".join(path)] = np.asarray(v)
        np.savez(os.path.join(checkpoint_path, f"params_{step}.npz"), **flat)
        with open(os.path.join(checkpoint_path, "config.json"), "w") as f:
            json.dump(self.config, f)
        if self.dataset_statistics is not None:
            with open(os.path.join(checkpoint_path, "dataset_statistics.json"), "w") as f:
                json.dump(self.dataset_statistics, f, default=lambda x: np.asarray(x).tolist())

    @property
    def head_spec(self) -> "M.HeadSpec":
        """README mix head (one readout token) or DiscreteActionHead on 4 / 28 readout tokens (base_network.py:22-33)."""
        return M.HeadSpec.from_config(self.config)

    @property
    def runtime(self):
        if self._runtime is None:
            from .runtime import Runtime
            self._runtime = Runtime(self.params, self.precision, self.device, self.head_spec)
        elif self._runtime._params_id != id(self.params):
            # ``model.params = ema_params`` (the reference's EMA swap is base_model.replace(params=...), data/simpler/
            # evaluate.py:443): re-upload the device blobs and drop the captured graphs that point into the old ones
            self._runtime.upload(self.params)
        return self._runtime

    def replace(self, **changes) -> "HyperVLA":
        """flax.struct.dataclass.replace of the reference (evaluate.py:443 ``model.replace(params=...)``): a new model object;
        the device runtime is rebuilt on first use when ``params`` changed."""
        import dataclasses
        new = dataclasses.replace(self, **changes)
        if any(k in changes and changes[k] is not getattr(self, k) for k in ("params", "precision", "device")):
            new._runtime = None
        return new

    # ---- initial-image encoder --------------------------------------------------------------
    def set_initial_image_encoder(self, dino_tree: Optional[dict]) -> None:
        """Use a separate frozen DINOv2-base param tree (HF Flax names) for the initial image, as the reference's eval
        loop does with the pretrained facebook/dinov2-base (data/simpler/evaluate.py:146-163, scripts/train.py:187-194).
        ``None`` (default) re-uses the model's own shared image_encoder leaves."""
        self._init_encoder_blobs = None if dino_tree is None else self.runtime.pack_dino_tree(dino_tree)

    def encode_initial_image(self, images):
        """DINO_encode_image of the reference (data/simpler/evaluate.py:155-163) on the GPU: uint8 (T,224,224,3) host or
        CUDA images -> ``initial_state`` dict whose ``patch_embeddings`` (T,257,768) stay on the device, so a task switch
        (encode -> create_tasks) never leaves the GPU."""
        import torch
        rt = self.runtime
        img = images if torch.is_tensor(images) else torch.from_numpy(np.ascontiguousarray(images))
        if img.dim() == 5 and img.shape[1] == 1:
            img = img[:, 0]
        img = img.to(rt.device)
        hidden = rt.dino_forward(img, getattr(self, "_init_encoder_blobs", None))
        return {"image_primary": images, "patch_embeddings": hidden.float(),
                "pad_mask_dict": {"image_primary": np.ones((int(img.shape[0]), 1))}}

    # ---- generate ------------------------------------------------------------------------------------
    def create_tasks(self, goals=None, instruction_dict: dict = None, initial_state=None, *, task_ids=None, base_params=None):
        """Build the ``tasks`` dict and generate base-net parameters (reference: model.py:35-83).
        Returns ``(base_params, tasks, intermediate_states)``.

        Task-switch scheduler (batched form of the per-episode reset, data/utils/hypervla_interface.py:141-146,
        data/simpler/evaluate.py:263-277): with ``task_ids`` (k distinct slots) and the ``base_params`` returned by an
        earlier call, only those k rows are regenerated, IN PLACE, from the k instructions / initial states given; the
        other rows are untouched bit for bit and the weight buffer keeps its address, so the captured act graph is reused."""
        if instruction_dict is None or "language_instruction" not in instruction_dict:
            raise ValueError("create_tasks needs instruction_dict['language_instruction']")
        lang = instruction_dict["language_instruction"]
        batch_size = int(np.asarray(lang["input_ids"]).shape[0]) if "input_ids" in lang else int(np.shape(lang["token_embedding"])[0])
        tasks = {"pad_mask_dict": {}}
        if self.example_batch is not None and "task" in self.example_batch:
            for k, v in self.example_batch["task"].items():
                if k not in ("pad_mask_dict", "language_instruction"):
                    v = np.asarray(v)
                    tasks[k] = np.zeros((batch_size, *v.shape[1:]), dtype=v.dtype)
        for k in list(tasks.keys()):
            if k != "pad_mask_dict":
                tasks["pad_mask_dict"][k] = np.zeros(batch_size, dtype=bool)
        tasks["pad_mask_dict"]["language_instruction"] = np.ones(batch_size, dtype=bool)
        tasks["language_instruction"] = lang
        if (task_ids is None) != (base_params is None):
            raise ValueError("task_ids and base_params go together (regenerate rows of an existing GeneratedBaseParams in place)")
        if base_params is not None and not isinstance(base_params, GeneratedBaseParams):
            raise TypeError("base_params must be the object returned by HyperVLA.create_tasks")
        (base_params, _ctx), intermediate_states = self.hypernet.apply(
            {"params": self.params}, tasks, train=False, initial_states=initial_state, model=self,
            task_ids=task_ids, base_params=base_params)
        return base_params, tasks, intermediate_states

    # ---- act -------------------------------------------------------------------------------------------
    def sample_actions(self, images, instruction_dict=None, task=None, timestep_pad_mask=None, base_params=None,
                       train: bool = False, rng=None, image_embeddings=None, *, task_index=None, return_attention_maps: bool = False):
        """One control step (reference: model.py:85-137).  ``images`` (B, 1, 224, 224, 3) uint8 -- host
        array (numpy / pinned torch) or CUDA tensor.  Host in -> numpy out (includes the device->host read
        the caller performs at hypervla_interface.py:207); CUDA in -> CUDA tensors out, asynchronous.
        Returns ``(action (B,4,7) float32, intermediate_states)``.

        ``intermediate_states`` is ``{"gripper_logits": (B,4)}``; with ``return_attention_maps=True`` (a debugging call, off the
        captured-graph path) it also carries the tree the reference sows (model.py:125-137, read by InferenceWrapper at
        data/utils/hypervla_interface.py:208-217): ``["intermediates"]["encoder"]["DINO_attention_map"][0]`` = tuple of 12 arrays
        (B,12,257,257) and ``[...]["Transformer_0"]["encoderblock_i"]["MultiHeadDotProductAttention_0"]["attention_weights"][0]``
        = (B,4,257,257)."""
        if train:
            raise ValueError("hvla is an inference path: train=True is unsupported")
        if image_embeddings is not None:
            raise ValueError("image_embeddings are only used by the Siglip encoder, which is outside the supported config")
        if not isinstance(base_params, GeneratedBaseParams):
            raise TypeError("base_params must be the object returned by HyperVLA.create_tasks")
        ti = task_index if task_index is not None else base_params.task_index
        if self.head_spec.kind == "discrete":
            if return_attention_maps:
                raise ValueError("attention maps are produced for the mix-head configuration")
            action, tokens = self.base_net.apply({"params": base_params}, images, None, timestep_pad_mask, rng=rng, train=False,
                                                 method=BaseNetwork.predict_action, model=self, task_index=ti)
            return action, {"action_tokens": tokens}
        if return_attention_maps:
            action, logits, dmaps, bmaps = self.base_net.apply({"params": base_params}, images, None, timestep_pad_mask, rng=rng, train=False,
                                                               method=BaseNetwork.predict_action, model=self, task_index=ti,
                                                               attention_maps=True)
            enc = {"DINO_attention_map": (tuple(dmaps[l] for l in range(Cfg.DINO_LAYERS)),),
                   "Transformer_0": {f"encoderblock_{i}": {"MultiHeadDotProductAttention_0": {"attention_weights": (bmaps[i],)}}
                                     for i in range(Cfg.BASE_LAYERS)}}
            return action, {"gripper_logits": logits, "intermediates": {"encoder": enc}}
        action, logits = self.base_net.apply({"params": base_params}, images, None, timestep_pad_mask, rng=rng, train=False,
                                             method=BaseNetwork.predict_action, model=self, task_index=ti)
        return action, {"gripper_logits": logits}
