#!/usr/bin/env python
"""Phase timestamps of the tcgen05 attention kernel (CTA 0): HVLA_ATTN_TS=1 python tools/attn_ts.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
from hvla import _native as N  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = N.lib()
st = int(torch.cuda.current_stream().cuda_stream)
qkv = (torch.randn(B * 257, 2304, device="cuda") * 0.5).to(torch.bfloat16)
out = torch.empty(B * 257, 768, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    N.check(lib.hvla_dino_attention(st, qkv.data_ptr(), out.data_ptr(), B, 1), "attn")
    torch.cuda.synchronize()
    print("----", file=sys.stderr)
