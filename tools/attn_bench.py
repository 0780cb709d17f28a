#!/usr/bin/env python
"""Times the DINOv2 attention kernels alone (CUDA events): python tools/attn_bench.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
from hvla import _native as N  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = N.lib()
st = int(torch.cuda.current_stream().cuda_stream)
qkv = torch.randn(B * 257, 2304, device="cuda")
qkv[:, :768] *= 0.35
qkv = qkv.to(torch.bfloat16)
out = torch.empty(B * 257, 768, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for impl in (0, 1):
    for _ in range(3):
        N.check(lib.hvla_dino_attention(st, qkv.data_ptr(), out.data_ptr(), B, impl), "attn")
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.hvla_dino_attention(st, qkv.data_ptr(), out.data_ptr(), B, impl)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    fl = B * 12 * 4 * 257 * 257 * 64
    print(f"impl {impl} B={B}: {ts[len(ts)//2]*1e3:8.1f} us  {fl / ts[len(ts)//2] / 1e9:6.1f} TFLOP/s")
