"""Stand-in for `transformers.FlaxDinov2Model` (transformers==4.50.0, un-vendored, dropped from the
transformers 5.x on this box) built on the *torch* `transformers.Dinov2Model` that is installed here:
same library, same published DINOv2-base architecture, weights carried over from the HF-Flax-named tree.

TEST INFRASTRUCTURE (see oracle/refshim/__init__.py).  The one step HF-torch and HF-Flax do differently —
resampling the 37x37 position grid to 16x16 (`jax.image.resize(..., 'bicubic', antialias=False)` vs
`torch.nn.functional.interpolate(mode='bicubic')`, different cubic coefficient) — is done with the
restatement in oracle/hypervla_oracle.py:interpolate_pos_table and handed to torch as a ready 257-row
table, so that step stays UNPINNED (weights-only host preprocessing)."""
from __future__ import annotations

import numpy as np

from . import flaxlite as nn
from . import jaxlite as J

HIDDEN, LAYERS, HEADS, MLP, PATCH, POS_ROWS = 768, 12, 12, 3072, 14, 1370


def random_flax_tree(seed=0, dtype=np.float32):
    """A parameter tree with the HF-Flax DINOv2 names/shapes."""
    rng = np.random.default_rng(seed)
    n = lambda *s: (rng.standard_normal(s, dtype=np.float32) * 0.02).astype(dtype)
    dense = lambda i, o: {"kernel": n(i, o), "bias": n(o)}
    ln = lambda: {"scale": np.ones(HIDDEN, dtype), "bias": np.zeros(HIDDEN, dtype)}
    layer = lambda: {
        "norm1": ln(), "norm2": ln(),
        "attention": {"attention": {k: dense(HIDDEN, HIDDEN) for k in ("query", "key", "value")},
                      "output": {"dense": dense(HIDDEN, HIDDEN)}},
        "layer_scale1": {"lambda1": np.ones(HIDDEN, dtype)}, "layer_scale2": {"lambda1": np.ones(HIDDEN, dtype)},
        "mlp": {"fc1": dense(HIDDEN, MLP), "fc2": dense(MLP, HIDDEN)}}
    return {
        "embeddings": {"cls_token": n(1, 1, HIDDEN), "mask_token": np.zeros((1, HIDDEN), dtype),
                       "position_embeddings": n(1, POS_ROWS, HIDDEN),
                       "patch_embeddings": {"projection": {"kernel": n(PATCH, PATCH, 3, HIDDEN), "bias": n(HIDDEN)}}},
        "encoder": {"layer": {str(i): layer() for i in range(LAYERS)}},
        "layernorm": ln()}


def _load_torch(model, tree, pos_table):
    import torch
    sd = model.state_dict()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float64)))
    emb = tree["embeddings"]
    sd["embeddings.cls_token"].copy_(t(emb["cls_token"]))
    sd["embeddings.position_embeddings"].copy_(t(pos_table)[None])
    proj = emb["patch_embeddings"]["projection"]
    sd["embeddings.patch_embeddings.projection.weight"].copy_(t(np.asarray(proj["kernel"]).transpose(3, 2, 0, 1)))
    sd["embeddings.patch_embeddings.projection.bias"].copy_(t(proj["bias"]))
    for l in range(LAYERS):
        L, p = tree["encoder"]["layer"][str(l)], f"encoder.layer.{l}."
        for nm in ("query", "key", "value"):
            sd[p + f"attention.attention.{nm}.weight"].copy_(t(np.asarray(L["attention"]["attention"][nm]["kernel"]).T))
            sd[p + f"attention.attention.{nm}.bias"].copy_(t(L["attention"]["attention"][nm]["bias"]))
        sd[p + "attention.output.dense.weight"].copy_(t(np.asarray(L["attention"]["output"]["dense"]["kernel"]).T))
        sd[p + "attention.output.dense.bias"].copy_(t(L["attention"]["output"]["dense"]["bias"]))
        for i in ("1", "2"):
            sd[p + f"layer_scale{i}.lambda1"].copy_(t(L[f"layer_scale{i}"]["lambda1"]))
            sd[p + f"norm{i}.weight"].copy_(t(L[f"norm{i}"]["scale"]))
            sd[p + f"norm{i}.bias"].copy_(t(L[f"norm{i}"]["bias"]))
            sd[p + f"mlp.fc{i}.weight"].copy_(t(np.asarray(L["mlp"][f"fc{i}"]["kernel"]).T))
            sd[p + f"mlp.fc{i}.bias"].copy_(t(L["mlp"][f"fc{i}"]["bias"]))
    sd["layernorm.weight"].copy_(t(tree["layernorm"]["scale"]))
    sd["layernorm.bias"].copy_(t(tree["layernorm"]["bias"]))
    model.load_state_dict(sd)


_MODEL_CACHE = {}


class _Outputs:
    def __init__(self, last_hidden_state, attentions):
        self.last_hidden_state = last_hidden_state
        self.attentions = attentions


class FlaxDinov2Module(nn.Module):
    """`FlaxDinov2Model(config).module`: NHWC pixels in, `.last_hidden_state` (B,257,768) out."""
    config: object = None

    def __call__(self, pixel_values, output_attentions=False, **_unused):
        import torch
        import transformers
        from oracle import hypervla_oracle as O

        tree = self.params_subtree()
        if tree is None:
            assert self.is_initializing(), "image_encoder parameters missing"
            tree = random_flax_tree(0)
            self.set_params_subtree(tree)
        px = np.asarray(pixel_values, np.float64)
        assert px.shape[1:] == (224, 224, 3), px.shape
        cfg = transformers.Dinov2Config(image_size=224, patch_size=PATCH)
        assert (cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.mlp_ratio, cfg.hidden_act,
                cfg.layer_norm_eps, cfg.qkv_bias, cfg.use_swiglu_ffn) == (HIDDEN, LAYERS, HEADS, 4, "gelu", 1e-6, True, False)
        try:
            cfg._attn_implementation = "eager"
        except Exception:
            pass
        key = (np.asarray(tree["embeddings"]["cls_token"]).tobytes(),
               np.asarray(tree["encoder"]["layer"]["11"]["mlp"]["fc2"]["bias"]).tobytes())
        model = _MODEL_CACHE.get(key)
        with torch.no_grad():
            if model is None:
                model = transformers.Dinov2Model(cfg).double().eval()
                pos = O.interpolate_pos_table(np.asarray(tree["embeddings"]["position_embeddings"], np.float32))
                _load_torch(model, tree, pos)
                _MODEL_CACHE.clear()
                _MODEL_CACHE[key] = model
            out = model(pixel_values=torch.from_numpy(px.transpose(0, 3, 1, 2).copy()), output_attentions=bool(output_attentions))
        att = getattr(out, "attentions", None)
        att = tuple(J.wrap(a.numpy()) for a in att) if att else None
        return _Outputs(J.wrap(out.last_hidden_state.numpy()), att)


class FlaxDinov2Model:
    """`FlaxDinov2Model.from_pretrained(...)` / `FlaxDinov2Model(config).module` (base_vit.py:74-77)."""

    def __init__(self, config=None, **_kwargs):
        self.config = config
        self.module = FlaxDinov2Module(config)
        self._params = None

    @property
    def params(self):
        if self._params is None:
            self._params = random_flax_tree(1)
        return self._params

    @classmethod
    def from_pretrained(cls, name, **_kwargs):
        import transformers
        assert name == "facebook/dinov2-base", name
        return cls(transformers.Dinov2Config(image_size=518, patch_size=PATCH))
