"""The supported configuration envelope of the B200 hot path.

The values are the README configuration of the reference
(/root/reference/README.md:18-63 on top of the defaults in
scripts/configs/hypervla_pretrain_config.py:326-400).  The CUDA kernels are
compiled for exactly these shapes; anything else raises ``ValueError`` before
any launch (there is no CPU fallback and no generic path).
"""
from __future__ import annotations

import copy

# ---- shape constants the kernels are specialised for --------------------------
IMAGE_SIZE = 224
PATCH = 14
GRID = IMAGE_SIZE // PATCH          # 16
N_PATCH = GRID * GRID               # 256
DINO_TOKENS = N_PATCH + 1           # 257 (CLS + patches)
DINO_DIM = 768
DINO_LAYERS = 12
DINO_HEADS = 12
DINO_HEAD_DIM = 64
DINO_MLP = 3072
DINO_POS_GRID = 37                  # dinov2-base: image_size 518 / 14
DINO_PATCH_K = PATCH * PATCH * 3    # 588

CTX_DIM = 128
CTX_LAYERS = 6
CTX_HEADS = 4
CTX_MLP = 512
LANG_TOKENS = 32
LANG_DIM = 768
CTX_TOKENS = LANG_TOKENS + 1 + 1    # language + initial-image CLS + layer token

BASE_DIM = 64
BASE_LAYERS = 4
BASE_HEADS = 4
BASE_MLP = 128
BASE_TOKENS = N_PATCH + 1           # 256 patches + 1 action token
ACTION_HORIZON = 4
ACTION_DIM = 7

DINO_IMAGE_MEAN = (0.485, 0.456, 0.406)
DINO_IMAGE_STD = (0.229, 0.224, 0.225)


def default_config() -> dict:
    """README configuration as the plain dict ``config.json`` would hold."""
    return {
        "seed": 2025,
        "window_size": 1,
        "model": {},
        "hypernet_kwargs": dict(
            encoder_type="transformer",
            context_embedding_dim=CTX_DIM,
            context_encoder_kwargs=dict(
                num_layers=CTX_LAYERS,
                mlp_dim=CTX_MLP,
                num_attention_heads=CTX_HEADS,
                dropout_rate=0.0,
                attention_dropout_rate=0.0,
                add_position_embedding=False,
            ),
            attend_to_padding=False,
            task_attend_to_layer=False,
            embedding_dropout_rate=0.0,
            scale_context_embedding=True,
            one_hot_context=False,
            output_head_bias=True,
            generation_strategy="block",
            shared_modules=("image_encoder",),
            include_goal_image=False,
            use_initial_image=True,
            use_all_image_tokens=False,
            share_TF_output_head=False,
            init_strategy=0,
            share_all_params=False,
            share_layer_index=True,
            image_dropout=0.0,
        ),
        "base_net_kwargs": dict(
            model_type="vit",
            action_head_type="mix",
            action_horizon=ACTION_HORIZON,
            action_dim=ACTION_DIM,
            cnn_kwargs={},
            vit_kwargs=dict(
                encoder_type="DINOv2",
                patch_size=16,
                hidden_dim=BASE_DIM,
                num_layers=BASE_LAYERS,
                num_heads=BASE_HEADS,
                mlp_dim=BASE_MLP,
                dropout_rate=0.0,
                use_language_token=False,
                fine_tune_pretrained_image_encoder=True,
                image_embedding_noise=0.0,
                use_differential_transformer=False,
                return_attention_map=False,
                add_positional_embedding=True,
                include_class_token=False,
            ),
            action_head_kwargs=dict(
                token_per_horizon=False,
                squash_continuous_action=True,
                tanh_scaling_factor=5.0,
                clip_target=True,
                max_action=5.0,
                hidden_dims=(),
            ),
        ),
    }


def _require(cond: bool, what: str) -> None:
    if not cond:
        raise ValueError(
            f"hvla: unsupported configuration ({what}); the sm_100a kernels are built "
            "for the README HyperVLA configuration only and there is no fallback path"
        )


def validate_config(config: dict) -> dict:
    """Check ``config`` against the supported envelope; return a normalised copy."""
    cfg = copy.deepcopy(config)
    hk = cfg.get("hypernet_kwargs")
    bk = cfg.get("base_net_kwargs")
    _require(isinstance(hk, dict) and isinstance(bk, dict), "missing hypernet_kwargs/base_net_kwargs")
    if "action_head_kwargs" not in bk:  # same default the reference injects (model.py:157-163)
        bk["action_head_kwargs"] = dict(token_per_horizon=False, squash_continuous_action=True,
                                        clip_target=False, max_action=5.0)
    ce = hk.get("context_encoder_kwargs", {})
    vk = bk.get("vit_kwargs", {})
    ak = bk["action_head_kwargs"]
    _require(hk.get("context_embedding_dim") == CTX_DIM, "context_embedding_dim != 128")
    _require(ce.get("num_layers") == CTX_LAYERS, "context encoder num_layers != 6")
    _require(ce.get("mlp_dim") == CTX_MLP, "context encoder mlp_dim != 512")
    _require(ce.get("num_attention_heads") == CTX_HEADS, "context encoder heads != 4")
    _require(not ce.get("add_position_embedding", False), "context add_position_embedding")
    _require(hk.get("generation_strategy") == "block", "generation_strategy != 'block'")
    _require(bool(hk.get("share_layer_index")), "share_layer_index must be True")
    _require(tuple(hk.get("shared_modules", ())) == ("image_encoder",), "shared_modules != ('image_encoder',)")
    _require(bool(hk.get("use_initial_image")), "use_initial_image must be True")
    _require(not hk.get("use_all_image_tokens", False), "use_all_image_tokens")
    _require(bool(hk.get("scale_context_embedding")), "scale_context_embedding must be True")
    _require(not hk.get("attend_to_padding", False), "attend_to_padding")
    _require(not hk.get("task_attend_to_layer", False), "task_attend_to_layer")
    _require(not hk.get("include_goal_image", False), "include_goal_image")
    _require(not hk.get("share_TF_output_head", False), "share_TF_output_head")
    _require(not hk.get("share_all_params", False), "share_all_params")
    _require(hk.get("output_head_bias", True), "output_head_bias=False")
    _require(bk.get("model_type") == "vit", "model_type != 'vit'")
    _require(bk.get("action_head_type") in ("mix", "discrete"), "action_head_type not in ('mix', 'discrete')")
    _require(bk.get("action_horizon") == ACTION_HORIZON and bk.get("action_dim") == ACTION_DIM,
             "action_horizon/action_dim != 4/7")
    _require(vk.get("encoder_type") == "DINOv2", "encoder_type != 'DINOv2'")
    _require(vk.get("hidden_dim") == BASE_DIM and vk.get("num_layers") == BASE_LAYERS
             and vk.get("num_heads") == BASE_HEADS and vk.get("mlp_dim") == BASE_MLP,
             "base ViT is not 4L/64d/4h/mlp128")
    _require(not vk.get("use_language_token", False), "use_language_token")
    _require(vk.get("add_positional_embedding", True), "add_positional_embedding=False")
    _require(not vk.get("include_class_token", False), "include_class_token")
    _require(not vk.get("use_differential_transformer", False), "use_differential_transformer")
    _require(not vk.get("return_attention_map", False), "return_attention_map")
    _require(float(vk.get("image_embedding_noise", 0.0)) == 0.0, "image_embedding_noise > 0")
    if bk.get("action_head_type") == "discrete":
        # DiscreteActionHead (action_heads.py:252-396) on 4 or 28 readout tokens, 256 uniform bins over [-1, 1] (its defaults)
        _require(ak.get("discrete_token_type") in ("action_horizon", "action_dim_and_action_horizon"),
                 "discrete_token_type not in ('action_horizon', 'action_dim_and_action_horizon')")
        return cfg
    _require(not ak.get("token_per_horizon", False), "token_per_horizon")
    _require(ak.get("squash_continuous_action", True), "squash_continuous_action=False")
    _require(len(tuple(ak.get("hidden_dims", ()))) == 0, "action head hidden_dims")
    # the head epilogues compute tanh(x / 5) * 5 (MixActionHead defaults, action_heads.py:439, 445); a checkpoint trained with
    # another tanh_scaling_factor / max_action (e.g. max_action=1 for bounds normalisation) must not load silently
    _require(float(ak.get("tanh_scaling_factor", 5.0)) == 5.0, "tanh_scaling_factor != 5")
    _require(float(ak.get("max_action", 5.0)) == 5.0, "max_action != 5")
    return cfg
