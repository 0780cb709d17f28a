"""Hypernetwork parameter pytree: synthetic init, and packing into device blobs.

Tree names/shapes follow the reference (`SURVEY.md` Appendix A.1):
``task_token_projection``, ``initial_image_projection``, ``*_pos_embedding``,
``context_encoder``, ``output_head_<leafpath>`` x73 (hypernetwork.py:43-67, 80-86)
and ``encoder_image_encoder_<leafpath>`` x223 flat vectors (hypernetwork.py:88-97,
model.py:344).

There are no checkpoints in this environment (no network), so ``init_params``
draws a synthetic tree from ``numpy.random.default_rng(seed)``:
  * variant "P0": reference-faithful init -- output-head kernels are ZERO and head
    biases hold a base-net init draw (BIAS_INIT, hypernetwork.py:72-77, model.py:328-346);
  * variant "P1": trained-like -- P0 with non-zero head kernels, perturbed LayerNorm
    scales/biases, so the context encoder and head GEMM actually matter.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

from . import config as C
from . import metadata as M

F32 = np.float32


# ----------------------------------------------------------------------------------
# init helpers
# ----------------------------------------------------------------------------------
def _xavier(rng, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(F32)


def _normal(rng, shape, std):
    return (rng.standard_normal(size=shape) * std).astype(F32)


def _trunc_normal(rng, shape, std):
    x = rng.standard_normal(size=shape)
    return (np.clip(x, -2.0, 2.0) * std).astype(F32)


def _init_transformer(rng, d, layers, heads, mlp, variant):
    hd = d // heads
    tree = {}
    for l in range(layers):
        blk = {}
        for ln in ("LayerNorm_0", "LayerNorm_1"):
            if variant == "P1":
                blk[ln] = {"scale": (1.0 + _normal(rng, (d,), 0.1)), "bias": _normal(rng, (d,), 0.02)}
            else:
                blk[ln] = {"scale": np.ones((d,), F32), "bias": np.zeros((d,), F32)}
        att = {}
        for nm in ("query", "key", "value"):
            att[nm] = {"kernel": _xavier(rng, (d, heads, hd), d, d),
                       "bias": _normal(rng, (heads, hd), 0.02) if variant == "P1" else np.zeros((heads, hd), F32)}
        att["out"] = {"kernel": _xavier(rng, (heads, hd, d), d, d),
                      "bias": _normal(rng, (d,), 0.02) if variant == "P1" else np.zeros((d,), F32)}
        blk["MultiHeadDotProductAttention_0"] = att
        bstd = 0.02 if variant == "P1" else 1e-6
        blk["MlpBlock_0"] = {
            "Dense_0": {"kernel": _xavier(rng, (d, mlp), d, mlp), "bias": _normal(rng, (mlp,), bstd)},
            "Dense_1": {"kernel": _xavier(rng, (mlp, d), mlp, d), "bias": _normal(rng, (d,), bstd)},
        }
        tree[f"encoderblock_{l}"] = blk
    if variant == "P1":
        tree["encoder_norm"] = {"scale": 1.0 + _normal(rng, (d,), 0.1), "bias": _normal(rng, (d,), 0.02)}
    else:
        tree["encoder_norm"] = {"scale": np.ones((d,), F32), "bias": np.zeros((d,), F32)}
    return tree


def _init_dinov2(rng, variant):
    """Synthetic facebook/dinov2-base-shaped tree (no pretrained weights offline)."""
    D, Fm = C.DINO_DIM, C.DINO_MLP
    qk_gain = 3.0 if variant == "P1" else 1.0   # peakier attention in the trained-like set

    def dense(i, o, gain=1.0):
        return {"kernel": _trunc_normal(rng, (i, o), 0.02 * gain),
                "bias": _normal(rng, (o,), 0.02) if variant == "P1" else np.zeros((o,), F32)}

    def ln():
        if variant == "P1":
            return {"scale": 1.0 + _normal(rng, (D,), 0.1), "bias": _normal(rng, (D,), 0.02)}
        return {"scale": np.ones((D,), F32), "bias": np.zeros((D,), F32)}

    layers = {}
    for i in range(C.DINO_LAYERS):
        layers[str(i)] = {
            "norm1": ln(),
            "attention": {
                "attention": {"query": dense(D, D, qk_gain), "key": dense(D, D, qk_gain), "value": dense(D, D)},
                "output": {"dense": dense(D, D)},
            },
            "layer_scale1": {"lambda1": (np.ones((D,), F32) if variant == "P0"
                                         else (1.0 + _normal(rng, (D,), 0.1)))},
            "norm2": ln(),
            "mlp": {"fc1": dense(D, Fm), "fc2": dense(Fm, D)},
            "layer_scale2": {"lambda1": (np.ones((D,), F32) if variant == "P0"
                                         else (1.0 + _normal(rng, (D,), 0.1)))},
        }
    n_pos = C.DINO_POS_GRID ** 2 + 1
    return {
        "embeddings": {
            "cls_token": _trunc_normal(rng, (1, 1, D), 0.02),
            "mask_token": np.zeros((1, D), F32),
            "position_embeddings": _trunc_normal(rng, (1, n_pos, D), 0.02),
            "patch_embeddings": {"projection": {
                "kernel": _trunc_normal(rng, (C.PATCH, C.PATCH, 3, D), 0.02),
                "bias": _normal(rng, (D,), 0.02) if variant == "P1" else np.zeros((D,), F32)}},
        },
        "encoder": {"layer": layers},
        "layernorm": ln(),
    }


def init_base_params(rng, variant="P0", spec: "M.HeadSpec" = M.MIX) -> dict:
    """One ``BaseNetwork.init`` draw (the values BIAS_INIT copies into head biases).  The draw order of the mix head is frozen
    (the committed golden fixtures depend on it)."""
    d = C.BASE_DIM
    n_cont = C.ACTION_HORIZON * (C.ACTION_DIM - 1)
    enc = {
        "image_encoder": _init_dinov2(rng, variant),
        "image_embedding_projection": {"kernel": _xavier(rng, (C.DINO_DIM, d), C.DINO_DIM, d),
                                       "bias": np.zeros((d,), F32)},
        "pos_embedding": _normal(rng, (1, spec.tokens, d), 0.02),
        "Transformer_0": _init_transformer(rng, d, C.BASE_LAYERS, C.BASE_HEADS, C.BASE_MLP, "P0"),
    }
    if spec.kind == "mix":
        head = {
            "continuous_head": {"kernel": _xavier(rng, (d, n_cont), d, n_cont), "bias": np.zeros((n_cont,), F32)},
            "discrete_head": {"kernel": _xavier(rng, (d, C.ACTION_HORIZON), d, C.ACTION_HORIZON),
                              "bias": np.zeros((C.ACTION_HORIZON,), F32)},
        }
    else:
        v = spec.vocab_out
        head = {"vocab_proj": {"kernel": _xavier(rng, (d, v), d, v), "bias": np.zeros((v,), F32)}}
    return {"encoder": enc, "action_head": head}


def init_params(seed: int = 2025, variant: str = "P1", spec: "M.HeadSpec" = M.MIX) -> dict:
    """Hypernetwork param pytree (what ``HyperVLA.params`` holds)."""
    if variant not in ("P0", "P1"):
        raise ValueError(f"unknown params variant {variant!r}")
    rng = np.random.default_rng(seed)
    d = C.CTX_DIM
    p = {
        "task_token_projection": {"kernel": _xavier(rng, (C.LANG_DIM, d), C.LANG_DIM, d),
                                  "bias": _normal(rng, (d,), 0.02) if variant == "P1" else np.zeros((d,), F32)},
        "initial_image_projection": {"kernel": _xavier(rng, (C.DINO_DIM, d), C.DINO_DIM, d),
                                     "bias": _normal(rng, (d,), 0.02) if variant == "P1" else np.zeros((d,), F32)},
        "task_pos_embedding": _normal(rng, (1, C.LANG_TOKENS, d), 0.02),
        "initial_image_pos_embedding": _normal(rng, (1, 1, d), 0.02),
        "layer_pos_embedding": _normal(rng, (1, 1, d), 0.02),
        "context_encoder": _init_transformer(rng, d, C.CTX_LAYERS, C.CTX_HEADS, C.CTX_MLP, variant),
    }
    base = init_base_params(rng, variant, spec)
    for path, value in M.iter_leaves(base):
        name = M.head_name(path)
        if M.is_generated(path):
            n = value.size
            if variant == "P1":
                kern = _normal(rng, (d, n), 0.05)
            else:
                kern = np.zeros((d, n), F32)
            p[f"output_head_{name}"] = {"kernel": kern, "bias": value.ravel().copy()}
        else:
            p[name] = value.ravel().copy()          # shared leaf: flat vector (model.py:344)
    return p


# ----------------------------------------------------------------------------------
# DINOv2 position-table interpolation (weights-only preprocessing, done once on host)
# ----------------------------------------------------------------------------------
def _keys_cubic(x):
    out = ((1.5 * x - 2.5) * x) * x + 1.0
    out = np.where(x >= 1.0, ((-0.5 * x + 2.5) * x - 4.0) * x + 2.0, out)
    return np.where(x >= 2.0, 0.0, out)


def _resize_weights(n_in: int, n_out: int, scale: np.float32) -> np.ndarray:
    """Weight matrix [n_in, n_out] of jax.image.scale_and_translate(method='bicubic',
    antialias=False, translation=0): Keys cubic a=-0.5, renormalised over in-range taps."""
    inv = F32(1.0) / F32(scale)
    sample = (np.arange(n_out, dtype=F32) + F32(0.5)) * inv - F32(0.5)
    x = np.abs(sample[None, :] - np.arange(n_in, dtype=F32)[:, None])
    w = _keys_cubic(x).astype(F32)
    tot = w.sum(axis=0, keepdims=True)
    w = np.where(np.abs(tot) > 1000.0 * np.finfo(F32).eps, w / np.where(tot != 0, tot, 1), 0).astype(F32)
    ok = (sample >= -0.5) & (sample <= n_in - 0.5)
    return np.where(ok[None, :], w, 0).astype(F32)


def interpolate_pos_table(position_embeddings: np.ndarray) -> np.ndarray:
    """(1,1370,768) -> (257,768): what HF FlaxDinov2 ``interpolate_pos_encoding`` yields
    for a 224x224 input (37x37 -> 16x16, the ``+0.1`` size trick).
    The un-vendored transformers==4.50.0 source is not on this box: PARITY UNPINNED."""
    pe = np.asarray(position_embeddings, F32)[0]
    cls_pe, patch_pe = pe[:1], pe[1:]
    g = C.DINO_POS_GRID
    patch_pe = patch_pe.reshape(g, g, C.DINO_DIM)
    scale = F32((C.GRID + 0.1) / g)
    w = _resize_weights(g, C.GRID, scale)                      # [37,16]
    out = np.einsum("hwc,ha,wb->abc", patch_pe, w, w, optimize=True).astype(F32)
    return np.concatenate([cls_pe, out.reshape(C.N_PATCH, C.DINO_DIM)], axis=0)


# ----------------------------------------------------------------------------------
# views
# ----------------------------------------------------------------------------------
def dino_tree_from_params(params: dict) -> dict:
    """Rebuild the HF DINOv2 tree from the ``encoder_image_encoder_*`` flat vectors."""
    tree: dict = {}
    for path, shape in M.shared_leaves():
        vec = params[M.head_name(path)]
        M.set_path(tree, path[2:], np.asarray(vec, F32).reshape(shape))   # strip ('encoder','image_encoder')
    return tree


# ----------------------------------------------------------------------------------
# packing into the blobs the C ABI takes (layouts documented in include/hvla.h)
# ----------------------------------------------------------------------------------
def _cat(chunks):
    return np.ascontiguousarray(np.concatenate([np.asarray(c, F32).ravel() for c in chunks]))


def pack_hn_blob(params: dict) -> np.ndarray:
    """Context-encoder side of the hypernetwork as one fp32 vector (hvla.h: HN blob)."""
    d, Hh = C.CTX_DIM, C.CTX_HEADS
    ch = [params["task_token_projection"]["kernel"], params["task_token_projection"]["bias"],
          params["initial_image_projection"]["kernel"], params["initial_image_projection"]["bias"],
          params["task_pos_embedding"], params["initial_image_pos_embedding"], params["layer_pos_embedding"]]
    enc = params["context_encoder"]
    for l in range(C.CTX_LAYERS):
        b = enc[f"encoderblock_{l}"]
        a = b["MultiHeadDotProductAttention_0"]
        wqkv = np.concatenate([a[n]["kernel"].reshape(d, d) for n in ("query", "key", "value")], axis=1)
        bqkv = np.concatenate([a[n]["bias"].reshape(d) for n in ("query", "key", "value")])
        ch += [b["LayerNorm_0"]["scale"], b["LayerNorm_0"]["bias"], wqkv, bqkv,
               a["out"]["kernel"].reshape(d, d), a["out"]["bias"],
               b["LayerNorm_1"]["scale"], b["LayerNorm_1"]["bias"],
               b["MlpBlock_0"]["Dense_0"]["kernel"], b["MlpBlock_0"]["Dense_0"]["bias"],
               b["MlpBlock_0"]["Dense_1"]["kernel"], b["MlpBlock_0"]["Dense_1"]["bias"]]
    ch += [enc["encoder_norm"]["scale"], enc["encoder_norm"]["bias"]]
    blob = _cat(ch)
    assert blob.size == hn_blob_size(), (blob.size, hn_blob_size())
    return blob


def hn_blob_size() -> int:
    d, m = C.CTX_DIM, C.CTX_MLP
    per_layer = 2 * d + d * 3 * d + 3 * d + d * d + d + 2 * d + d * m + m + m * d + d
    return 2 * (C.LANG_DIM * d + d) + (C.LANG_TOKENS + 2) * d + C.CTX_LAYERS * per_layer + 2 * d


def pack_heads(params: dict, spec: "M.HeadSpec" = M.MIX) -> Tuple[np.ndarray, np.ndarray]:
    """Output heads as ONE matrix: W [128, NGP] and b [NGP] in packed column order
    (the 73 ``Dense(128->leaf)`` heads of hypernetwork.py:65-67 side by side)."""
    NGP = M.n_generated_padded(spec)
    W = np.zeros((C.CTX_DIM, NGP), F32)
    b = np.zeros((NGP,), F32)
    for path, (off, shape) in M.packed_offsets(spec).items():
        n = int(np.prod(shape))
        head = params[f"output_head_{M.head_name(path)}"]
        W[:, off:off + n] = head["kernel"]
        b[off:off + n] = head["bias"]
    return W, b


def dino_vec_layout() -> Dict[str, Tuple[int, int]]:
    """name -> (offset, size) inside the fp32 DINOv2 vector blob (hvla.h: DINO vec blob)."""
    D, Fm = C.DINO_DIM, C.DINO_MLP
    lay, off = {}, 0

    def add(name, n):
        nonlocal off
        lay[name] = (off, n)
        off += n

    add("patch_b", D)
    add("cls", D)
    add("pos", C.DINO_TOKENS * D)
    for l in range(C.DINO_LAYERS):
        for nm, n in (("ln1_s", D), ("ln1_b", D), ("bqkv", 3 * D), ("bo", D), ("ls1", D),
                      ("ln2_s", D), ("ln2_b", D), ("b1", Fm), ("b2", D), ("ls2", D),
                      # LayerNorm folded into the following linear layer (bf16 tensor-core path, see pack_dino_tree)
                      ("bqkv_f", 3 * D), ("cs_qkv", 3 * D), ("b1_f", Fm), ("cs_1", Fm)):
            add(f"l{l}.{nm}", n)
    add("lnf_s", D)
    add("lnf_b", D)
    add("pos_blk", C.N_PATCH * D)      # position rows of the patch tokens in the blocked stream layout (csrc/gemm_tc.cuh: xblk_f4)
    lay["__total__"] = (off, 0)
    return lay


DINO_PATCH_K_PAD = 640      # im2col K (588) padded to a multiple of the 64-wide bf16 K tile


def dino_mat_layout(transposed: bool) -> Dict[str, Tuple[int, Tuple[int, int]]]:
    """name -> (offset, (rows, cols)) inside the DINOv2 matrix blob.  ``transposed`` is the
    bf16 tensor-core layout ([N,K], K contiguous); otherwise Flax [K,N]."""
    D, Fm, KP = C.DINO_DIM, C.DINO_MLP, DINO_PATCH_K_PAD
    lay, off = {}, 0

    def add(name, k, n):
        nonlocal off
        lay[name] = (off, (n, k) if transposed else (k, n))
        off += k * n

    add("patch_w", KP, D)
    for l in range(C.DINO_LAYERS):
        add(f"l{l}.wqkv", D, 3 * D)
        add(f"l{l}.wo", D, D)
        add(f"l{l}.w1", D, Fm)
        add(f"l{l}.w2", Fm, D)
    lay["__total__"] = (off, (0, 0))
    return lay


def pack_dino(params: dict, transposed: bool) -> Tuple[np.ndarray, np.ndarray]:
    """(vec blob fp32, matrix blob fp32) of the shared DINOv2 leaves held in the hypernetwork params."""
    return pack_dino_tree(dino_tree_from_params(params), transposed)


def split_matrices_x3(mat_kn: np.ndarray) -> np.ndarray:
    """DINOv2 matrix blob of the split-operand flow (csrc/dino_x3.cuh, dtype HVLA_BF16X3) from the fp32 Flax [K,N] blob: every matrix
    becomes [N, 3K] = [hi | hi | lo] of its transpose, hi = bf16(w), lo = bf16(w - hi), at 3x its element offset.  Returned as the
    uint16 bit patterns of the bf16 values (view as torch.bfloat16)."""
    lay = dino_mat_layout(False)
    out = np.empty(3 * lay["__total__"][0], np.uint16)
    for name, (off, shape) in lay.items():
        if name == "__total__":
            continue
        k, n = shape
        wt = np.ascontiguousarray(mat_kn[off:off + k * n].reshape(k, n).T, F32)       # [N, K]
        hi = bf16_round(wt)
        lo = bf16_round(wt - hi)
        hb, lb = (hi.view(np.uint32) >> 16).astype(np.uint16), (lo.view(np.uint32) >> 16).astype(np.uint16)
        out[3 * off:3 * off + 3 * k * n] = np.concatenate([hb, hb, lb], axis=1).ravel()
    return out


def bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest bfloat16 (ties to even), returned as float32: what the tensor cores will multiply."""
    u = np.ascontiguousarray(x, F32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(F32).reshape(np.shape(x))


def pack_dino_tree(t: dict, transposed: bool) -> Tuple[np.ndarray, np.ndarray]:
    """(vec blob fp32, matrix blob fp32) of a HF DINOv2-base param tree.  The caller casts the matrix blob to
    bf16 for the tensor-core path.  The position table is interpolated here, once.

    Tensor-core layout (``transposed``): the LayerNorm in front of the q|k|v and fc1 linears is folded into them,
    ``LN(x) W + b = xhat (gamma*W) + (beta W + b)`` with ``xhat = (x - mean) rstd``: the matrix blob holds ``gamma*W``, the
    vector blob the folded biases ``bqkv_f`` / ``b1_f`` and the column sums ``cs_*`` of the bf16-rounded ``gamma*W`` (for
    kernels that apply mean / rstd after the matrix product: ``rstd (x W' - mean cs) + b_f``).  The query third of
    ``gamma*Wqkv`` and ``bqkv_f`` also carries the 1/sqrt(64) query scaling."""
    D = C.DINO_DIM
    vl, ml = dino_vec_layout(), dino_mat_layout(transposed)
    vec = np.zeros((vl["__total__"][0],), F32)
    mat = np.zeros((ml["__total__"][0],), F32)

    def putv(name, arr):
        o, n = vl[name]
        vec[o:o + n] = np.asarray(arr, F32).ravel()

    def putm(name, w_kn):
        o, (r, c) = ml[name]
        w = np.asarray(w_kn, F32)
        mat[o:o + r * c] = (w.T if transposed else w).ravel()

    emb = t["embeddings"]
    putv("patch_b", emb["patch_embeddings"]["projection"]["bias"])
    putv("cls", emb["cls_token"])
    pos = np.asarray(interpolate_pos_table(emb["position_embeddings"]), F32).reshape(C.DINO_TOKENS, D)
    putv("pos", pos)
    # rows 1..256 again, blocked like the fp32 residual stream of the large-batch flow: [p >> 5][col >> 2][p & 31][col & 3]
    putv("pos_blk", pos[1:].reshape(C.N_PATCH // 32, 32, D // 4, 4).transpose(0, 2, 1, 3))
    wp = np.zeros((DINO_PATCH_K_PAD, D), F32)
    wp[:C.DINO_PATCH_K] = emb["patch_embeddings"]["projection"]["kernel"].reshape(C.DINO_PATCH_K, D)
    putm("patch_w", wp)
    for l in range(C.DINO_LAYERS):
        L = t["encoder"]["layer"][str(l)]
        a = L["attention"]["attention"]
        putv(f"l{l}.ln1_s", L["norm1"]["scale"]); putv(f"l{l}.ln1_b", L["norm1"]["bias"])
        putv(f"l{l}.bqkv", np.concatenate([a["query"]["bias"], a["key"]["bias"], a["value"]["bias"]]))
        putv(f"l{l}.bo", L["attention"]["output"]["dense"]["bias"])
        putv(f"l{l}.ls1", L["layer_scale1"]["lambda1"])
        putv(f"l{l}.ln2_s", L["norm2"]["scale"]); putv(f"l{l}.ln2_b", L["norm2"]["bias"])
        putv(f"l{l}.b1", L["mlp"]["fc1"]["bias"]); putv(f"l{l}.b2", L["mlp"]["fc2"]["bias"])
        putv(f"l{l}.ls2", L["layer_scale2"]["lambda1"])
        wqkv = np.concatenate([a["query"]["kernel"], a["key"]["kernel"], a["value"]["kernel"]], axis=1).astype(F32)
        w1 = np.asarray(L["mlp"]["fc1"]["kernel"], F32)
        for nm, w, g, be, b in (("qkv", wqkv, L["norm1"]["scale"], L["norm1"]["bias"], vec[vl[f"l{l}.bqkv"][0]:][:3 * D]),
                                ("1", w1, L["norm2"]["scale"], L["norm2"]["bias"], np.asarray(L["mlp"]["fc1"]["bias"], F32))):
            wf = np.asarray(g, F32)[:, None] * w
            bf = (np.asarray(be, np.float64) @ w.astype(np.float64) + b).astype(F32)
            if nm == "qkv":                # q / sqrt(64) (flax scales the query before QK^T): a power of two, exact in bf16
                wf[:, :D] *= F32(0.125)
                bf[:D] *= F32(0.125)
            putv(f"l{l}.b{nm}_f", bf)
            putv(f"l{l}.cs_{nm}", bf16_round(wf).astype(np.float64).sum(0).astype(F32))
            if transposed:
                putm(f"l{l}.w{nm}", wf)
            else:
                putm(f"l{l}.w{nm}", w)
        putm(f"l{l}.wo", L["attention"]["output"]["dense"]["kernel"])
        putm(f"l{l}.w2", L["mlp"]["fc2"]["kernel"])
    putv("lnf_s", t["layernorm"]["scale"]); putv("lnf_b", t["layernorm"]["bias"])
    return vec, mat


def unpack_generated(row: np.ndarray, spec: "M.HeadSpec" = M.MIX) -> dict:
    """Packed per-task weight row -> base-net pytree (generated leaves only, Flax names)."""
    tree: dict = {}
    row = np.asarray(row)
    for path, (off, shape) in M.packed_offsets(spec).items():
        n = int(np.prod(shape))
        M.set_path(tree, path, row[..., off:off + n].reshape(row.shape[:-1] + tuple(shape)))
    return tree
