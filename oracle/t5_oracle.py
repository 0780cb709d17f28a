"""CPU oracle of the T5-base token embedder that feeds the hypernetwork (TEST INFRASTRUCTURE ONLY).

The reference embeds the tokenised instruction with `FlaxT5EncoderModel(config).module` (t5-base) and passes
`last_hidden_state` as `token_embedding` (octo/model/components/tokenizers.py:186-211, data/utils/language_tokenizer.py:9-28,
data/simpler/evaluate.py:240-262).  The arithmetic lives in un-vendored transformers==4.50.0 (Flax T5); it is restated here
from the published T5 v1.0 encoder and PINNED against the torch `transformers.T5EncoderModel` of the local transformers
(same library, same architecture) in tests/test_t5.py:
  x = shared[input_ids]
  12 x { h = rms(x)*w0 ; q,k,v = h Wq, h Wk, h Wv (no bias) ; s = q k^T (NOT scaled) + rel_bias + (1-mask)*finfo.min ;
         x += softmax(s) v Wo ; h = rms(x)*w1 ; x += relu(h Wi) Wo2 }
  out = rms(x)*wf          rms(x) = x * rsqrt(mean(x^2) + 1e-6)   (no mean subtraction, no bias)
  rel_bias[h,i,j] = table[bucket(j - i), h], bidirectional, 32 buckets, max distance 128, taken from block 0 for all blocks.
Weights use the HF torch layout (Linear.weight = [out, in]).
"""
import numpy as np

D, H, HD, FF, LAYERS, BUCKETS, MAXDIST, EPS = 768, 12, 64, 3072, 12, 32, 128, 1e-6


def relative_bucket(rel):
    """T5Attention._relative_position_bucket(bidirectional=True, num_buckets=32, max_distance=128); rel = key - query."""
    nb = BUCKETS // 2
    out = (rel > 0).astype(np.int64) * nb
    n = np.abs(rel)
    max_exact = nb // 2
    is_small = n < max_exact
    with np.errstate(divide="ignore"):
        large = max_exact + (np.log(np.maximum(n, 1).astype(np.float32) / max_exact) / np.log(MAXDIST / max_exact) * (nb - max_exact)).astype(np.int64)
    large = np.minimum(large, nb - 1)
    return out + np.where(is_small, n, large)


def rms(x, w):
    dt = x.dtype
    var = np.mean(x.astype(dt) ** 2, axis=-1, keepdims=True)
    return x * (1.0 / np.sqrt(var + dt.type(EPS))) * w.astype(dt)


def encode(sd, input_ids, attention_mask, dtype=np.float32):
    """sd: HF torch-style state dict of T5EncoderModel as numpy arrays -> last_hidden_state (T,S,768)."""
    dt = np.dtype(dtype)
    g = lambda k: np.asarray(sd[k]).astype(dt)
    ids = np.asarray(input_ids)
    T, S = ids.shape
    x = g("shared.weight")[ids]
    pos = np.arange(S)
    bucket = relative_bucket(pos[None, :] - pos[:, None])                                   # [query, key]
    bias = g("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight")[bucket].transpose(2, 0, 1)   # (H,S,S)
    neg = (1.0 - np.asarray(attention_mask).astype(dt)) * np.finfo(dt).min                     # (T,S)
    for l in range(LAYERS):
        p = f"encoder.block.{l}.layer."
        h = rms(x, g(p + "0.layer_norm.weight"))
        q = (h @ g(p + "0.SelfAttention.q.weight").T).reshape(T, S, H, HD)
        k = (h @ g(p + "0.SelfAttention.k.weight").T).reshape(T, S, H, HD)
        v = (h @ g(p + "0.SelfAttention.v.weight").T).reshape(T, S, H, HD)
        s = np.einsum("tqhd,tkhd->thqk", q, k) + bias[None] + neg[:, None, None, :]
        s = s - s.max(-1, keepdims=True)
        e = np.exp(s)
        a = e / e.sum(-1, keepdims=True)
        o = np.einsum("thqk,tkhd->tqhd", a, v).reshape(T, S, D)
        x = x + o @ g(p + "0.SelfAttention.o.weight").T
        h = rms(x, g(p + "1.layer_norm.weight"))
        x = x + np.maximum(h @ g(p + "1.DenseReluDense.wi.weight").T, 0) @ g(p + "1.DenseReluDense.wo.weight").T
    return rms(x, g("encoder.final_layer_norm.weight"))
