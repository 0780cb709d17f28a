"""CPU oracle for the HyperVLA inference hot path (generate -> act).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`hyper-vla_b200/`) may
import this module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs do, and only as the checker / CPU baseline.

It is a plain-NumPy restatement of the reference's JAX/Flax forward, function by
function, each citing the reference file:line it follows (paths relative to
/root/reference).  `dtype=np.float32` mirrors the reference's JAX-on-CPU numerics
(true fp32 matmuls); `dtype=np.float64` is the ground truth used to bound the fp32
rounding noise of both this oracle and the CUDA path.

PARITY STATUS
  * The reference ships no tests, fixtures or golden vectors and JAX/Flax
    are not installable here (F3).  The oracle is pinned instead against the reference's
    OWN hot-path code executed in this container through `oracle/refshim/` (NumPy
    stand-ins for the jax/flax primitives only): committed fixtures
    tests/golden/ref_*.npz (generator: tests/golden/make_ref_golden.py), checked in
    tests/test_reference_pin.py.  Pinned that way: hypernetwork.py, transformer.py,
    base_vit.py, base_network.py, action_heads.py (MixActionHead), model.py
    (from_config / init_base_net / create_tasks / sample_actions).
  * Restated rather than executed (flax 0.8.1 / jax sources are not on this box): Dense,
    LayerNorm, MultiHeadDotProductAttention, gelu — once in refshim/flaxlite.py for the
    reference run and once below for the oracle.
  * The DINOv2 arithmetic lives in un-vendored `transformers==4.50.0`
    (`FlaxDinov2Module`, requirements_full_install.txt:22).  Its block stack is pinned
    against the *torch* `transformers.Dinov2Model` of the local transformers 5.5.0
    (same published architecture) in tests/test_oracle_dinov2_torch.py; only the
    bicubic position-table interpolation (`jax.image.scale_and_translate`) remains
    **unpinned** and is isolated in `interpolate_pos_table`.
"""
from __future__ import annotations

import math

import numpy as np

# ---- model constants (README config) ---------------------------
PATCH, GRID, N_PATCH = 14, 16, 256
DINO_DIM, DINO_LAYERS, DINO_HEADS, DINO_POS_GRID = 768, 12, 12, 37
CTX_DIM, CTX_LAYERS, CTX_HEADS = 128, 6, 4
BASE_DIM, BASE_LAYERS, BASE_HEADS = 64, 4, 4
ACTION_HORIZON, ACTION_DIM = 4, 7
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)


# =====================================================================================
# Flax primitives (flax==0.8.1 documented behaviour)
# =====================================================================================
def layer_norm(x, scale, bias, eps=1e-6):
    """flax.linen.LayerNorm, use_fast_variance=True: var = max(0, E[x^2] - E[x]^2);
    y = (x - mean) * (rsqrt(var + eps) * scale) + bias."""
    dt = x.dtype
    mean = x.mean(axis=-1, keepdims=True, dtype=dt)
    mean2 = (x * x).mean(axis=-1, keepdims=True, dtype=dt)
    var = np.maximum(dt.type(0), mean2 - mean * mean)
    mul = (dt.type(1) / np.sqrt(var + dt.type(eps))) * scale.astype(dt)
    return (x - mean) * mul + bias.astype(dt)


def gelu_tanh(x):
    """flax.linen.gelu default (approximate=True) -- transformer.py:66."""
    dt = x.dtype
    c = dt.type(math.sqrt(2.0 / math.pi))
    return dt.type(0.5) * x * (dt.type(1) + np.tanh(c * (x + dt.type(0.044715) * (x * x * x))))


def gelu_erf(x):
    """Exact GELU (HF ACT2FN['gelu']) used inside DINOv2."""
    dt = x.dtype
    from scipy.special import erf
    return dt.type(0.5) * x * (dt.type(1) + erf(x * dt.type(1.0 / math.sqrt(2.0))))


def softmax(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


_ATTENTION_SINK = None      # list collecting every attention-weight tensor (N,H,S,S) in call order, see capture_attention()


class capture_attention:
    """``with capture_attention() as maps: sample_actions(...)`` -> maps = the softmax attention weights of every attention call
    in order: the 12 DINOv2 layers (what FlaxDinov2 returns as ``outputs.attentions``, sown at base_vit.py:118), then the 4 base
    encoder blocks (``sow_weights`` / ``attention_map``, transformer.py:172-191) -- the reference's ``intermediates``."""

    def __enter__(self):
        global _ATTENTION_SINK
        _ATTENTION_SINK = []
        return _ATTENTION_SINK

    def __exit__(self, *exc):
        global _ATTENTION_SINK
        _ATTENTION_SINK = None
        return False


def _attend(q, k, v, mask):
    """q,k,v (N,S,H,e) -> (N,S,H*e): q/sqrt(e), masked softmax(qk^T) v (batched matmuls, BLAS)."""
    dt = q.dtype
    q = q / dt.type(math.sqrt(q.shape[-1]))
    s = np.matmul(q.transpose(0, 2, 1, 3), k.transpose(0, 2, 3, 1))          # (N,H,S,S)
    if mask is not None:
        s = np.where(mask, s, np.finfo(dt).min)
    w = softmax(s)
    if _ATTENTION_SINK is not None:
        _ATTENTION_SINK.append(w.copy())
    o = np.matmul(w, v.transpose(0, 2, 1, 3))                                 # (N,H,S,e)
    N, H, S, e = o.shape
    return o.transpose(0, 2, 1, 3).reshape(N, S, H * e)


def mha(x, p, mask):
    """flax.linen.MultiHeadDotProductAttention as called at transformer.py:183-190:
    q,k,v DenseGeneral -> q / sqrt(head_dim) -> where(mask, s, finfo.min) -> softmax ->
    weighted sum -> DenseGeneral over (heads, head_dim).  x: (N,S,d); mask: (N,1,S,S) bool."""
    dt = x.dtype
    N, S, d = x.shape

    def proj(name):
        w = p[name]["kernel"].astype(dt)
        H, e = w.shape[1], w.shape[2]
        return (x @ w.reshape(d, H * e) + p[name]["bias"].astype(dt).reshape(H * e)).reshape(N, S, H, e)

    o = _attend(proj("query"), proj("key"), proj("value"), mask)
    wo = p["out"]["kernel"].astype(dt)
    return o @ wo.reshape(-1, wo.shape[-1]) + p["out"]["bias"].astype(dt)


def mha_per_sample(x, p, mask):
    """Same as `mha` but every sample has its own weights (leaves carry a leading N):
    the jax.vmap of scripts/train.py:559-579."""
    dt = x.dtype
    N, S, d = x.shape

    def proj(name):
        w = p[name]["kernel"].astype(dt)
        H, e = w.shape[2], w.shape[3]
        return (np.matmul(x, w.reshape(N, d, H * e)) + p[name]["bias"].astype(dt).reshape(N, 1, H * e)).reshape(N, S, H, e)

    o = _attend(proj("query"), proj("key"), proj("value"), mask)
    wo = p["out"]["kernel"].astype(dt)
    return np.matmul(o, wo.reshape(N, -1, wo.shape[-1])) + p["out"]["bias"].astype(dt)[:, None]


# =====================================================================================
# hypervla/components/transformer.py
# =====================================================================================
def encoder_block(x, p, mask, per_sample=False):
    """Encoder1DBlock (transformer.py:127-201): x + MHA(LN(x)); then + MLP(LN(.)), tanh-GELU."""
    dt = x.dtype
    if per_sample:
        def vec(a):
            return a.astype(dt)[:, None, :]
        y = (lambda s, b: _ln_ps(x, s, b))(p["LayerNorm_0"]["scale"], p["LayerNorm_0"]["bias"])
        x = x + mha_per_sample(y, p["MultiHeadDotProductAttention_0"], mask)
        y = _ln_ps(x, p["LayerNorm_1"]["scale"], p["LayerNorm_1"]["bias"])
        m = p["MlpBlock_0"]
        h = np.matmul(y, m["Dense_0"]["kernel"].astype(dt)) + vec(m["Dense_0"]["bias"])
        h = gelu_tanh(h)
        h = np.matmul(h, m["Dense_1"]["kernel"].astype(dt)) + vec(m["Dense_1"]["bias"])
        return x + h
    y = layer_norm(x, p["LayerNorm_0"]["scale"], p["LayerNorm_0"]["bias"])
    x = x + mha(y, p["MultiHeadDotProductAttention_0"], mask)
    y = layer_norm(x, p["LayerNorm_1"]["scale"], p["LayerNorm_1"]["bias"])
    m = p["MlpBlock_0"]                                  # MlpBlock, transformer.py:42-75
    h = y @ m["Dense_0"]["kernel"].astype(dt) + m["Dense_0"]["bias"].astype(dt)
    h = gelu_tanh(h)
    h = h @ m["Dense_1"]["kernel"].astype(dt) + m["Dense_1"]["bias"].astype(dt)
    return x + h


def _ln_ps(x, scale, bias, eps=1e-6):
    dt = x.dtype
    mean = x.mean(axis=-1, keepdims=True, dtype=dt)
    mean2 = (x * x).mean(axis=-1, keepdims=True, dtype=dt)
    var = np.maximum(dt.type(0), mean2 - mean * mean)
    mul = (dt.type(1) / np.sqrt(var + dt.type(eps))) * scale.astype(dt)[:, None, :]
    return (x - mean) * mul + bias.astype(dt)[:, None, :]


def transformer(x, p, mask, num_layers, per_sample=False):
    """Transformer (transformer.py:204-262), add_position_embedding=False; final encoder_norm."""
    for l in range(num_layers):
        x = encoder_block(x, p[f"encoderblock_{l}"], mask, per_sample)
    if per_sample:
        return _ln_ps(x, p["encoder_norm"]["scale"], p["encoder_norm"]["bias"])
    return layer_norm(x, p["encoder_norm"]["scale"], p["encoder_norm"]["bias"])


# =====================================================================================
# hypervla/components/hypernetwork.py
# =====================================================================================
def context_mask(attention_mask, lang_pad):
    """Block mask of hypernetwork.py:151-181 for 32 language + 1 image + 1 layer token.
    cols 0..31: token_mask & lang_pad (every row); col 32: 1; col 33: 1 only on row 33
    (task_attend_to_layer=False).  Returns (T,1,34,34) bool."""
    T, L = attention_mask.shape
    S = L + 2
    m = np.zeros((T, 1, S, S), bool)
    lang = attention_mask.astype(bool) & lang_pad.astype(bool)[:, None]
    m[:, 0, :, :L] = lang[:, None, :]
    m[:, 0, :, L] = True
    m[:, 0, S - 1, S - 1] = True
    return m


def generate_context_embedding(params, token_embedding, attention_mask, lang_pad, init_cls, dtype=np.float32):
    """HyperNetwork.generate_context_embedding (hypernetwork.py:99-197).
    token_embedding (T,32,768); attention_mask (T,32); lang_pad (T,); init_cls (T,768) =
    initial_states['patch_embeddings'][:, 0].  Returns (T,1,128)."""
    dt = np.dtype(dtype)
    tok = token_embedding.astype(dt) @ params["task_token_projection"]["kernel"].astype(dt) \
        + params["task_token_projection"]["bias"].astype(dt)                          # :112
    tok = tok + params["task_pos_embedding"].astype(dt)                                # :115
    img = init_cls.astype(dt)[:, None, :] @ params["initial_image_projection"]["kernel"].astype(dt) \
        + params["initial_image_projection"]["bias"].astype(dt)                       # :126
    img = img + params["initial_image_pos_embedding"].astype(dt)                       # :127
    T = tok.shape[0]
    layer = np.zeros((T, 1, CTX_DIM), dt) + params["layer_pos_embedding"].astype(dt)   # :144-145
    ctx = np.concatenate([tok, img, layer], axis=1)                                    # :128, :147
    mask = context_mask(attention_mask, lang_pad)
    out = transformer(ctx, params["context_encoder"], mask, CTX_LAYERS)                # :184-186
    emb = out[:, -1:]                                                                  # :188
    return emb / dt.type(math.sqrt(CTX_DIM))                                           # :191-192


def generate(params, token_embedding, attention_mask, init_cls, lang_pad=None, dtype=np.float32,
             generated_paths=None):
    """HyperNetwork.__call__ (hypernetwork.py:199-219): every generated leaf is
    Dense_leaf(context_embedding[:, 0]) reshaped to (T, *leaf_shape).  `generated_paths` is an
    iterable of (path_tuple, shape); defaults to all `output_head_*` entries found.
    Returns (dict path_tuple -> array (T,*shape), context_embedding (T,1,128))."""
    dt = np.dtype(dtype)
    T = token_embedding.shape[0]
    if lang_pad is None:
        lang_pad = np.ones((T,), bool)                                                 # model.py:69
    emb = generate_context_embedding(params, token_embedding, attention_mask, lang_pad, init_cls, dt)
    e = emb[:, 0]                                                                      # token index 0 (:205)
    out = {}
    for path, shape in generated_paths:
        head = params["output_head_" + "_".join(path)]
        w = e @ head["kernel"].astype(dt) + head["bias"].astype(dt)                    # :227
        out[path] = w.reshape((T,) + tuple(shape))                                     # :217
    return out, emb


def to_tree(flat):
    tree = {}
    for path, v in flat.items():
        t = tree
        for k in path[:-1]:
            t = t.setdefault(k, {})
        t[path[-1]] = v
    return tree


# =====================================================================================
# DINOv2 (transformers==4.50.0 FlaxDinov2Module, un-vendored; restated from the published
# architecture, cross-checked against torch Dinov2Model in tests)
# =====================================================================================
def _keys_cubic(x):
    out = ((1.5 * x - 2.5) * x) * x + 1.0
    out = np.where(x >= 1.0, ((-0.5 * x + 2.5) * x - 4.0) * x + 2.0, out)
    return np.where(x >= 2.0, 0.0, out)


def interpolate_pos_table(position_embeddings):
    """FlaxDinov2Embeddings.interpolate_pos_encoding for 224x224: 37x37 -> 16x16 through
    jax.image.scale_and_translate(method='bicubic', antialias=False), scale=(16.1/37).
    PARITY UNPINNED (neither transformers 4.50 Flax source nor jax is on this box)."""
    f32 = np.float32
    pe = np.asarray(position_embeddings, f32)[0]
    g = DINO_POS_GRID
    grid = pe[1:].reshape(g, g, -1)
    inv = f32(1.0) / f32((GRID + 0.1) / g)
    sample = (np.arange(GRID, dtype=f32) + f32(0.5)) * inv - f32(0.5)
    x = np.abs(sample[None, :] - np.arange(g, dtype=f32)[:, None])
    w = _keys_cubic(x).astype(f32)
    tot = w.sum(0, keepdims=True)
    w = np.where(np.abs(tot) > 1000.0 * np.finfo(f32).eps, w / np.where(tot != 0, tot, 1), 0).astype(f32)
    w = np.where(((sample >= -0.5) & (sample <= g - 0.5))[None, :], w, 0).astype(f32)
    out = np.zeros((GRID, GRID, grid.shape[-1]), f32)
    tmp = np.tensordot(w.T, grid, axes=(1, 0))           # (16, 37, C): rows resized
    out = np.tensordot(w.T, tmp, axes=(1, 1))            # (16[w], 16[h], C)
    out = out.transpose(1, 0, 2)
    return np.concatenate([pe[:1], out.reshape(N_PATCH, -1)], 0).astype(f32)


def dinov2_embed(dino, images_u8, pos_table, dtype=np.float32):
    """base_vit.py:111-114 (normalise) + FlaxDinov2Embeddings: VALID 14x14/14 conv on NHWC,
    prepend CLS, add interpolated position table.  images (B,224,224,3) u8 -> (B,257,768)."""
    dt = np.dtype(dtype)
    x = images_u8.astype(dt) / dt.type(255.0)
    x = (x - np.asarray(IMAGE_MEAN, dt)) / np.asarray(IMAGE_STD, dt)
    B = x.shape[0]
    # im2col: (B,16,14,16,14,3) -> (B,256,(kh,kw,c))
    x = x.reshape(B, GRID, PATCH, GRID, PATCH, 3).transpose(0, 1, 3, 2, 4, 5).reshape(B, N_PATCH, PATCH * PATCH * 3)
    proj = dino["embeddings"]["patch_embeddings"]["projection"]
    w = proj["kernel"].astype(dt).reshape(PATCH * PATCH * 3, DINO_DIM)
    patches = x @ w + proj["bias"].astype(dt)
    cls = np.broadcast_to(dino["embeddings"]["cls_token"].astype(dt), (B, 1, DINO_DIM))
    return np.concatenate([cls, patches], axis=1) + pos_table.astype(dt)[None]


def dinov2_layer(x, L):
    """FlaxDinov2Layer: h = x + ls1*Attn(LN1(x)); y = h + ls2*MLP(LN2(h)); erf-GELU."""
    dt = x.dtype
    B, S, D = x.shape
    a = L["attention"]["attention"]
    y = layer_norm(x, L["norm1"]["scale"], L["norm1"]["bias"])
    hd = D // DINO_HEADS
    q = (y @ a["query"]["kernel"].astype(dt) + a["query"]["bias"].astype(dt)).reshape(B, S, DINO_HEADS, hd)
    k = (y @ a["key"]["kernel"].astype(dt) + a["key"]["bias"].astype(dt)).reshape(B, S, DINO_HEADS, hd)
    v = (y @ a["value"]["kernel"].astype(dt) + a["value"]["bias"].astype(dt)).reshape(B, S, DINO_HEADS, hd)
    o = _attend(q, k, v, None)
    o = o @ L["attention"]["output"]["dense"]["kernel"].astype(dt) + L["attention"]["output"]["dense"]["bias"].astype(dt)
    x = x + o * L["layer_scale1"]["lambda1"].astype(dt)
    y = layer_norm(x, L["norm2"]["scale"], L["norm2"]["bias"])
    h = gelu_erf(y @ L["mlp"]["fc1"]["kernel"].astype(dt) + L["mlp"]["fc1"]["bias"].astype(dt))
    h = h @ L["mlp"]["fc2"]["kernel"].astype(dt) + L["mlp"]["fc2"]["bias"].astype(dt)
    return x + h * L["layer_scale2"]["lambda1"].astype(dt)


def dinov2_forward(dino, images_u8, dtype=np.float32, pos_table=None):
    """FlaxDinov2Module(...).last_hidden_state: (B,224,224,3) u8 -> (B,257,768)."""
    if pos_table is None:
        pos_table = interpolate_pos_table(dino["embeddings"]["position_embeddings"])
    x = dinov2_embed(dino, images_u8, pos_table, dtype)
    for l in range(DINO_LAYERS):
        x = dinov2_layer(x, dino["encoder"]["layer"][str(l)])
    return layer_norm(x, dino["layernorm"]["scale"], dino["layernorm"]["bias"])


# =====================================================================================
# hypervla/components/base_vit.py + base_network.py + action_heads.py (per-sample weights)
# =====================================================================================
def base_mask(n, S, n_act=1):
    """base_vit.py:209-214: all ones, except patches (rows :-n_act) cannot see the action token(s)."""
    m = np.ones((n, 1, S, S), bool)
    m[:, :, :-n_act, -n_act:] = False
    return m


def base_vit_forward(gen, image_embeddings, dtype=np.float32, n_act=1):
    """ViT.__call__ after DINOv2 (base_vit.py:130-133, 157, 182-226) with per-sample weights.
    gen: base-net pytree whose generated leaves carry a leading B; image_embeddings (B,256,768)
    = last_hidden_state[:, 1:].  Returns the action embedding (B,64) (mix head: one readout token) or, with ``n_act`` > 1
    (DiscreteActionHead: base_network.py:22-33), the readout-token embeddings (B,n_act,64)."""
    dt = np.dtype(dtype)
    enc = gen["encoder"]
    x = image_embeddings.astype(dt)
    B = x.shape[0]
    pk = enc["image_embedding_projection"]
    patches = np.matmul(x, pk["kernel"].astype(dt)) + pk["bias"].astype(dt)[:, None, :]                  # :130-133
    tok = np.concatenate([patches, np.zeros((B, n_act, BASE_DIM), dt)], axis=1)       # :182-183
    tok = tok + enc["pos_embedding"].astype(dt).reshape(B, N_PATCH + n_act, BASE_DIM)  # :204
    out = transformer(tok, enc["Transformer_0"], base_mask(B, N_PATCH + n_act, n_act), BASE_LAYERS, per_sample=True)
    return out[:, -1] if n_act == 1 else out[:, -n_act:]                               # :226


def discrete_head(gen, h):
    """DiscreteActionHead.__call__ / predict_action(argmax=True) (action_heads.py:300-333, 372-396) + BinTokenizer.decode with
    256 uniform bins over [-1, 1] (octo/model/components/tokenizers.py:235-275).  h: readout tokens (B,n_tok,64), n_tok = 4
    (one token per horizon step, 7*256 logits each) or 28 (one per (horizon, dim), 256 logits each).
    Returns (action (B,4,7) f32 = bin centres, tokens (B,4,7) i32, logits (B,4,7,256))."""
    dt = h.dtype
    vp = gen["action_head"]["vocab_proj"]
    logits = np.einsum("btd,bdn->btn", h, vp["kernel"].astype(dt)) + vp["bias"].astype(dt)[:, None, :]
    logits = logits.reshape(h.shape[0], ACTION_HORIZON, ACTION_DIM, 256)             # :323-329
    tokens = logits.argmax(-1).astype(np.int32)                                        # :386 (first maximum, like jnp.argmax)
    thresholds = np.linspace(-1.0, 1.0, 257, dtype=np.float32)                         # tokenizers.py:251
    centres = (thresholds[1:] + thresholds[:-1]) / np.float32(2)                       # :273
    return centres[tokens].astype(np.float32), tokens, logits


def sample_actions_discrete(dino, gen_tree, images_u8, n_act, dtype=np.float32, pos_table=None):
    """sample_actions for the DiscreteActionHead configuration: -> (action, tokens, logits, readout tokens)."""
    hidden = dinov2_forward(dino, images_u8, dtype, pos_table)
    h = base_vit_forward(gen_tree, hidden[:, 1:], dtype, n_act=n_act)
    act, tokens, logits = discrete_head(gen_tree, h)
    return act, tokens, logits, h


def mix_head(gen, h):
    """MixActionHead.__call__/predict_action (action_heads.py:455-470, 536-537).
    Returns (action (B,4,7) f32, gripper logits (B,4))."""
    dt = h.dtype
    ah = gen["action_head"]
    cont = np.einsum("bd,bdn->bn", h, ah["continuous_head"]["kernel"].astype(dt)) + ah["continuous_head"]["bias"].astype(dt)
    cont = cont.reshape(-1, ACTION_HORIZON, ACTION_DIM - 1)                            # (h a) -> h a
    cont = np.tanh(cont / dt.type(5.0)) * dt.type(5.0)                                 # :470
    logit = np.einsum("bd,bdn->bn", h, ah["discrete_head"]["kernel"].astype(dt)) + ah["discrete_head"]["bias"].astype(dt)
    grip = (logit >= 0).astype(dt)                                                     # :536
    act = np.concatenate([cont, grip[..., None]], axis=-1)                             # :537
    return act.astype(np.float32), logit


def take_tasks(gen_flat, task_index):
    """Per-env weights from per-task weights (the batched extension)."""
    return {p: v[task_index] for p, v in gen_flat.items()}


def sample_actions(dino, gen_tree, images_u8, dtype=np.float32, pos_table=None, return_all=False):
    """HyperVLA.sample_actions (model.py:85-137) -> BaseNetwork.predict_action
    (base_network.py:170-183).  images (B,224,224,3) u8; gen_tree leaves carry leading B."""
    hidden = dinov2_forward(dino, images_u8, dtype, pos_table)
    h = base_vit_forward(gen_tree, hidden[:, 1:], dtype)                               # base_vit.py:122
    act, logit = mix_head(gen_tree, h)
    if return_all:
        return act, logit, hidden, h
    return act, logit
