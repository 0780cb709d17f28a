"""Base-network parameter tree, generated-vs-shared rule and packed layouts.

Mirrors what ``HyperVLA.init_base_net`` derives from a Flax ``init``
(/root/reference/hypervla/model.py:390-515): which base-net leaves the
hypernetwork generates, the per-leaf head output dims, the layer-token index
and the flattened head names (``flatten_dict``, model.py:532-540).  The
reference learns the shapes by running ``base_net.init``; here they are
written down from the module definitions
(components/base_vit.py:130-226, components/transformer.py:127-262,
components/action_heads.py:419-428, HF FlaxDinov2 param tree).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterator, List, Tuple

import numpy as np

from . import config as C

Path = Tuple[str, ...]


# ----------------------------------------------------------------------------------
# parameter-shape trees
# ----------------------------------------------------------------------------------
def _encoder_block_shapes(d: int, heads: int, mlp: int) -> dict:
    hd = d // heads
    attn = {
        "query": {"kernel": (d, heads, hd), "bias": (heads, hd)},
        "key": {"kernel": (d, heads, hd), "bias": (heads, hd)},
        "value": {"kernel": (d, heads, hd), "bias": (heads, hd)},
        "out": {"kernel": (heads, hd, d), "bias": (d,)},
    }
    return {
        "LayerNorm_0": {"scale": (d,), "bias": (d,)},
        "MultiHeadDotProductAttention_0": attn,
        "LayerNorm_1": {"scale": (d,), "bias": (d,)},
        "MlpBlock_0": {
            "Dense_0": {"kernel": (d, mlp), "bias": (mlp,)},
            "Dense_1": {"kernel": (mlp, d), "bias": (d,)},
        },
    }


def transformer_shapes(d: int, layers: int, heads: int, mlp: int) -> dict:
    """``Transformer`` param tree (components/transformer.py:247-260)."""
    tree = {f"encoderblock_{i}": _encoder_block_shapes(d, heads, mlp) for i in range(layers)}
    tree["encoder_norm"] = {"scale": (d,), "bias": (d,)}
    return tree


def dinov2_shapes() -> dict:
    """HF ``FlaxDinov2Module`` param tree for facebook/dinov2-base (un-vendored
    transformers==4.50.0; names as recalled in SURVEY.md Appendix A.2)."""
    D, F = C.DINO_DIM, C.DINO_MLP
    layer = {
        "norm1": {"scale": (D,), "bias": (D,)},
        "attention": {
            "attention": {
                "query": {"kernel": (D, D), "bias": (D,)},
                "key": {"kernel": (D, D), "bias": (D,)},
                "value": {"kernel": (D, D), "bias": (D,)},
            },
            "output": {"dense": {"kernel": (D, D), "bias": (D,)}},
        },
        "layer_scale1": {"lambda1": (D,)},
        "norm2": {"scale": (D,), "bias": (D,)},
        "mlp": {"fc1": {"kernel": (D, F), "bias": (F,)}, "fc2": {"kernel": (F, D), "bias": (D,)}},
        "layer_scale2": {"lambda1": (D,)},
    }
    n_pos = C.DINO_POS_GRID * C.DINO_POS_GRID + 1
    return {
        "embeddings": {
            "cls_token": (1, 1, D),
            "mask_token": (1, D),
            "position_embeddings": (1, n_pos, D),
            "patch_embeddings": {"projection": {"kernel": (C.PATCH, C.PATCH, 3, D), "bias": (D,)}},
        },
        "encoder": {"layer": {str(i): _copy_tree(layer) for i in range(C.DINO_LAYERS)}},
        "layernorm": {"scale": (D,), "bias": (D,)},
    }


def _copy_tree(t):
    return {k: _copy_tree(v) if isinstance(v, dict) else v for k, v in t.items()}


class HeadSpec:
    """Which action head the base network carries and how many readout (action) tokens its ViT appends
    (hypervla/components/base_network.py:22-33): the README ``mix`` head reads ONE token; ``DiscreteActionHead`` reads
    ``action_horizon`` (4) tokens, each projected to action_dim * 256 logits, or ``action_horizon * action_dim`` (28) tokens, each
    projected to 256 logits (components/action_heads.py:252-300)."""

    def __init__(self, kind: str = "mix", n_action_tokens: int = 1):
        if kind == "mix" and n_action_tokens != 1:
            raise ValueError("the mix head reads one action token (token_per_horizon=False)")
        if kind == "discrete" and n_action_tokens not in (C.ACTION_HORIZON, C.ACTION_HORIZON * C.ACTION_DIM):
            raise ValueError("discrete head: 4 (action_horizon) or 28 (action_dim_and_action_horizon) action tokens")
        if kind not in ("mix", "discrete"):
            raise ValueError(f"unsupported action head {kind!r}")
        self.kind, self.n_action_tokens = kind, n_action_tokens

    @property
    def vocab_out(self) -> int:          # outputs of vocab_proj per token
        return 0 if self.kind == "mix" else (C.ACTION_DIM * 256 if self.n_action_tokens == C.ACTION_HORIZON else 256)

    @property
    def tokens(self) -> int:             # base-ViT sequence length
        return C.N_PATCH + self.n_action_tokens

    def __eq__(self, other):
        return isinstance(other, HeadSpec) and (self.kind, self.n_action_tokens) == (other.kind, other.n_action_tokens)

    def __hash__(self):
        return hash((self.kind, self.n_action_tokens))

    def __repr__(self):
        return f"HeadSpec({self.kind!r}, {self.n_action_tokens})"

    @classmethod
    def from_config(cls, config: dict) -> "HeadSpec":
        bk = config["base_net_kwargs"]
        if bk.get("action_head_type") == "discrete":
            tok = bk.get("action_head_kwargs", {}).get("discrete_token_type")
            n = {"action_horizon": C.ACTION_HORIZON, "action_dim_and_action_horizon": C.ACTION_HORIZON * C.ACTION_DIM}.get(tok)
            if n is None:
                raise ValueError(f"discrete_token_type {tok!r} is not supported")
            return cls("discrete", n)
        return cls("mix", 1)


MIX = HeadSpec("mix", 1)


def base_net_shapes(spec: HeadSpec = MIX) -> dict:
    """``BaseNetwork`` param tree (SURVEY.md Appendix A.2 for the README ``mix`` config)."""
    d = C.BASE_DIM
    if spec.kind == "mix":
        head = {
            "continuous_head": {"kernel": (d, C.ACTION_HORIZON * (C.ACTION_DIM - 1)),
                                "bias": (C.ACTION_HORIZON * (C.ACTION_DIM - 1),)},
            "discrete_head": {"kernel": (d, C.ACTION_HORIZON), "bias": (C.ACTION_HORIZON,)},
        }
    else:       # DiscreteActionHead.setup: self.vocab_proj = nn.Dense(final_layer_size)  (action_heads.py:281-300)
        head = {"vocab_proj": {"kernel": (d, spec.vocab_out), "bias": (spec.vocab_out,)}}
    return {
        "encoder": {
            "image_encoder": dinov2_shapes(),
            "image_embedding_projection": {"kernel": (C.DINO_DIM, d), "bias": (d,)},
            "pos_embedding": (1, spec.tokens, d),
            "Transformer_0": transformer_shapes(d, C.BASE_LAYERS, C.BASE_HEADS, C.BASE_MLP),
        },
        "action_head": head,
    }


# ----------------------------------------------------------------------------------
# tree utilities (jax semantics: dict keys are visited in sorted order)
# ----------------------------------------------------------------------------------
def iter_leaves(tree: dict, prefix: Path = ()) -> Iterator[Tuple[Path, object]]:
    """Yield ``(path, leaf)`` in jax ``tree_flatten`` order (sorted dict keys)."""
    for k in sorted(tree.keys()):
        v = tree[k]
        if isinstance(v, dict):
            yield from iter_leaves(v, prefix + (k,))
        else:
            yield prefix + (k,), v


def get_path(tree: dict, path: Path):
    for k in path:
        tree = tree[k]
    return tree


def set_path(tree: dict, path: Path, value) -> None:
    for k in path[:-1]:
        tree = tree.setdefault(k, {})
    tree[path[-1]] = value


def is_generated(path: Path, shared_modules=("image_encoder",)) -> bool:
    """The ``filter`` rule of model.py:439-446: a leaf is shared (not generated)
    iff any key on its path *contains* one of ``shared_modules``."""
    for module in shared_modules:
        for key in path:
            if module in key:
                return False
    return True


def head_name(path: Path) -> str:
    """Flattened output-head name (model.py:512, 532-540; hypernetwork.py:222)."""
    return "_".join(path)


def build_base_net_metadata(config: dict) -> dict:
    """The ``base_net_metadata`` dict of model.py:460-513 for the supported config."""
    shared = tuple(config["hypernet_kwargs"].get("shared_modules", ()))
    shapes = base_net_shapes(HeadSpec.from_config(config))
    param_shape, param_dim, token_index, generation_flag = {}, {}, {}, {}
    output_head_info = OrderedDict()
    total = 0
    for path, shape in iter_leaves(shapes):
        n = int(np.prod(shape))
        total += n
        gen = is_generated(path, shared)
        set_path(param_shape, path, np.array(shape))
        set_path(param_dim, path, n)
        set_path(token_index, path, 0)            # share_layer_index=True (model.py:400-402)
        set_path(generation_flag, path, gen)
        output_head_info[head_name(path)] = dict(
            output_dim=n, generation_flag=gen, init_strategy=0, init_variance=0.0)
    return {
        "token_index_dict": token_index,
        "block_num": 1,
        "param_shape": param_shape,
        "total_param_num": total,
        "param_dim": param_dim,
        "generation_flag": generation_flag,
        "layer_token_mask": np.array([True]),
        "output_head_info": output_head_info,
    }


# ----------------------------------------------------------------------------------
# generated leaves: canonical (jax) order and the packed kernel order
# ----------------------------------------------------------------------------------
def generated_leaves_canonical(spec: HeadSpec = MIX) -> List[Tuple[Path, Tuple[int, ...]]]:
    """The generated leaves (73 for the mix head) in jax sorted-key order."""
    return [(p, s) for p, s in iter_leaves(base_net_shapes(spec)) if is_generated(p)]


def _blk(l: int, *rest: str) -> Path:
    return ("encoder", "Transformer_0", f"encoderblock_{l}") + rest


def generated_leaves_packed(spec: HeadSpec = MIX) -> List[Tuple[Path, Tuple[int, ...]]]:
    """Leaf order of the packed per-task weight row the kernels consume
    (layer-streaming order: what the base-net kernel touches first comes first).
    Every leaf is stored flat in its Flax row-major shape."""
    shapes = base_net_shapes(spec)
    order: List[Path] = [
        ("encoder", "image_embedding_projection", "kernel"),
        ("encoder", "image_embedding_projection", "bias"),
        ("encoder", "pos_embedding"),
    ]
    A = "MultiHeadDotProductAttention_0"
    for l in range(C.BASE_LAYERS):
        order += [
            _blk(l, "LayerNorm_0", "scale"), _blk(l, "LayerNorm_0", "bias"),
            _blk(l, A, "query", "kernel"), _blk(l, A, "query", "bias"),
            _blk(l, A, "key", "kernel"), _blk(l, A, "key", "bias"),
            _blk(l, A, "value", "kernel"), _blk(l, A, "value", "bias"),
            _blk(l, A, "out", "kernel"), _blk(l, A, "out", "bias"),
            _blk(l, "LayerNorm_1", "scale"), _blk(l, "LayerNorm_1", "bias"),
            _blk(l, "MlpBlock_0", "Dense_0", "kernel"), _blk(l, "MlpBlock_0", "Dense_0", "bias"),
            _blk(l, "MlpBlock_0", "Dense_1", "kernel"), _blk(l, "MlpBlock_0", "Dense_1", "bias"),
        ]
    order += [
        ("encoder", "Transformer_0", "encoder_norm", "scale"),
        ("encoder", "Transformer_0", "encoder_norm", "bias"),
    ]
    if spec.kind == "mix":
        order += [
            ("action_head", "continuous_head", "kernel"),
            ("action_head", "continuous_head", "bias"),
            ("action_head", "discrete_head", "kernel"),
            ("action_head", "discrete_head", "bias"),
        ]
    else:
        order += [("action_head", "vocab_proj", "kernel"), ("action_head", "vocab_proj", "bias")]
    out = [(p, tuple(get_path(shapes, p))) for p in order]
    assert sorted(p for p, _ in out) == sorted(p for p, _ in generated_leaves_canonical(spec))
    return out


N_GENERATED = 201_500
N_GENERATED_PADDED = 201_504      # row stride of the packed blob (multiple of 32 elements)


def n_generated(spec: HeadSpec = MIX) -> int:
    """Generated parameters per task: 201,500 (mix), 316,352 (discrete, 4 tokens), 218,048 (discrete, 28 tokens)."""
    return sum(int(np.prod(s)) for _, s in generated_leaves_packed(spec))


def n_generated_padded(spec: HeadSpec = MIX) -> int:
    """Row stride of the packed blob: the next multiple of 32 elements (hvla_generated_row_stride / hvla_discrete_row_stride)."""
    return (n_generated(spec) + 31) // 32 * 32


def packed_offsets(spec: HeadSpec = MIX) -> "OrderedDict[Path, Tuple[int, Tuple[int, ...]]]":
    """path -> (element offset in the packed row, shape)."""
    table: "OrderedDict[Path, Tuple[int, Tuple[int, ...]]]" = OrderedDict()
    off = 0
    for path, shape in generated_leaves_packed(spec):
        table[path] = (off, shape)
        off += int(np.prod(shape))
    assert off == n_generated(spec) and (spec != MIX or off == N_GENERATED), off
    return table


def shared_leaves() -> List[Tuple[Path, Tuple[int, ...]]]:
    """The 223 shared DINOv2 leaves (paths are full base-net paths)."""
    return [(p, s) for p, s in iter_leaves(base_net_shapes()) if not is_generated(p)]
