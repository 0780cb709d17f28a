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
