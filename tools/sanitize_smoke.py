#!/usr/bin/env python
"""Tiny generate+act (bf16 and fp32) for compute-sanitizer runs: python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
from hvla import config as C, params as P, synthetic as S  # noqa: E402
from hvla.model import HyperVLA  # noqa: E402

params = P.init_params(2025, "P1")
for prec in ("bf16", "fp32"):
    m = HyperVLA.from_config(C.default_config(), precision=prec, params=params)
    m.runtime.use_graphs = False
    inp = S.make_inputs(3, 2, 2)
    bp, tasks, _ = m.create_tasks(instruction_dict=inp["instruction_dict"], initial_state=inp["initial_state"])
    a, _ = m.sample_actions(inp["images"], None, tasks, None, bp)
    print(prec, a[0, 0])
    del m

# the steps either side of the model call
from hvla.postprocess import BatchedActionPostprocessor  # noqa: E402
from hvla.preprocess import BatchedImagePreprocessor  # noqa: E402

frames = np.random.default_rng(0).integers(0, 256, (2, 97, 131, 3), dtype=np.uint8)
for crop in (False, True):
    print("resize crop", crop, BatchedImagePreprocessor(224, crop=crop)(frames).float().mean().item())
stats = {"mean": np.zeros(7), "std": np.ones(7), "mask": np.ones(7, bool)}
pp = BatchedActionPostprocessor(2, "google_robot", "normal", stats)
print("post", pp.step(a)[1].cpu().numpy()[0])

# the step upstream of generate: T5 token embedder, tensor-core (split operands) and fp32 paths
if os.environ.get("HVLA_SANITIZE_T5", "1") != "0":
    from hvla import t5 as T5  # noqa: E402
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
    ids = rng.integers(1, 32000, (9, 13))
    am = (np.arange(13)[None, :] < rng.integers(1, 14, (9, 1))).astype(np.int64)
    for prec in ("bf16x3", "fp32"):
        print("t5", prec, T5.T5TokenEmbedder(sd, precision=prec)(ids, am).float().abs().mean().item())
