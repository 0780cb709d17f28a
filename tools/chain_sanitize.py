#!/usr/bin/env python
"""Small flow-B forward (GEMM chain + tcgen05 attention) for compute-sanitizer runs, and a soak loop:
  compute-sanitizer --tool memcheck python tools/chain_sanitize.py 2          # batch 2, flow B forced
  python tools/chain_sanitize.py 64 200                                       # 200 forwards at 64 images, every one bit-identical to the first"""
import os
import sys

import numpy as np

os.environ.setdefault("HVLA_FUSED_LN", "1")          # flow B (blocked stream + chain) also below 25 images
os.environ.setdefault("HVLA_CHAIN", "1")             # ... and above ~290
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
import torch  # noqa: E402
from hvla import params as P  # noqa: E402
from hvla.runtime import Runtime  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rt = Runtime(P.init_params(2025, "P1"), precision="bf16", device="cuda:0")
img = torch.from_numpy(np.random.default_rng(7).integers(0, 256, size=(B, 224, 224, 3), dtype=np.uint8)).cuda()
first = rt.dino_forward(img).view(torch.int16).clone()
bad = 0
for i in range(reps - 1):
    bad += int(not torch.equal(rt.dino_forward(img).view(torch.int16), first))
torch.cuda.synchronize()
print(f"B={B} lag={os.environ.get('HVLA_CHAIN_LAG', 'default')}: {reps} forwards, {bad} differ from the first, finite {bool(torch.isfinite(first.view(torch.bfloat16).float()).all())}")
sys.exit(1 if bad else 0)
