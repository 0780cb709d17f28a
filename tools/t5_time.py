"""Time the GPU T5 token embedder (hvla_t5_encode) for T instructions of 32 tokens: python tools/t5_time.py [T ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hyper-vla_b200")]
from hvla import t5 as T5

rng = np.random.default_rng(0)
sd = {"shared.weight": rng.standard_normal((T5.VOCAB, T5.D), dtype=np.float32),
      "encoder.final_layer_norm.weight": np.ones(T5.D, np.float32),
      "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight": rng.standard_normal((32, 12), dtype=np.float32)}
for l in range(T5.LAYERS):
    p = f"encoder.block.{l}.layer."
    for n, shp in (("0.SelfAttention.q", (768, 768)), ("0.SelfAttention.k", (768, 768)), ("0.SelfAttention.v", (768, 768)),
                   ("0.SelfAttention.o", (768, 768)), ("1.DenseReluDense.wi", (3072, 768)), ("1.DenseReluDense.wo", (768, 3072))):
        sd[p + n + ".weight"] = rng.standard_normal(shp, dtype=np.float32) * 0.03
    sd[p + "0.layer_norm.weight"] = np.ones(768, np.float32)
    sd[p + "1.layer_norm.weight"] = np.ones(768, np.float32)
precs = [a for a in sys.argv[1:] if not a.isdigit()] or ["bf16x3"]
for prec in precs:
    emb = T5.T5TokenEmbedder(sd, precision=prec)
    for T in [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 64]:
        ids = rng.integers(1, 32000, (T, 32)); am = np.ones((T, 32), np.int64)
        ids_d, am_d = torch.from_numpy(ids).cuda().int(), torch.from_numpy(am).cuda().int()
        for _ in range(3):
            emb(ids_d, am_d)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); emb(ids_d, am_d); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"t5_encode[{prec}] T={T}: p50 {np.median(ts):.3f} ms")
