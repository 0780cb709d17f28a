#!/usr/bin/env python
"""Yardstick for the attention kernel: library attention (torch SDPA backends: cuDNN, flash, memory-efficient; flash_attn if importable) on
the same problem -- 64 images x 12 heads x 257 tokens x 64 -- looped alone for a few seconds (sustained clocks), beside hvla_dino_attention.
Libraries are called here as a yardstick only."""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
from hvla import _native as N  # noqa: E402

B, H, S, D = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 12, 257, 64
secs = 1.5
lib = N.lib()
st = int(torch.cuda.current_stream().cuda_stream)
qkv = torch.randn(B * S, 3 * H * D, device="cuda")
qkv[:, :H * D] *= 0.35
qkv = qkv.to(torch.bfloat16)
out = torch.empty(B * S, H * D, device="cuda", dtype=torch.bfloat16)
flop = B * H * 4.0 * S * S * D


def loop(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:
        print(f"{name:44s} unavailable: {type(e).__name__}: {str(e)[:80]}")
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, n = time.perf_counter(), 0
    a.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        if n % 200 == 0:
            torch.cuda.current_stream().synchronize()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / n
    print(f"{name:44s} {us:8.1f} us   {flop / us / 1e6:6.0f} TFLOP/s", flush=True)


loop("hvla_dino_attention (tcgen05, this repo)", lambda: lib.hvla_dino_attention(st, qkv.data_ptr(), out.data_ptr(), B, 1))
loop("hvla_dino_attention (mma.sync, round-1 kernel)", lambda: lib.hvla_dino_attention(st, qkv.data_ptr(), out.data_ptr(), B, 0))
q4 = qkv.view(B, S, 3, H, D)
q, k, v = (q4[:, :, i].transpose(1, 2).contiguous() for i in range(3))       # [B, H, S, D], contiguous: the layout the libraries prefer
from torch.nn.attention import SDPBackend, sdpa_kernel  # noqa: E402
for nm, be in (("cuDNN", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION), ("mem-efficient", SDPBackend.EFFICIENT_ATTENTION)):
    def run(be=be):
        with sdpa_kernel(be):
            return F.scaled_dot_product_attention(q, k, v, scale=1.0)
    loop(f"torch SDPA {nm} backend ([B,H,S,D] contiguous)", run)
try:
    from flash_attn import flash_attn_func
    qf, kf, vf = (q4[:, :, i].contiguous() for i in range(3))                 # [B, S, H, D]
    loop("flash_attn_func 2.x ([B,S,H,D])", lambda: flash_attn_func(qf, kf, vf, softmax_scale=1.0))
except Exception as e:
    print("flash_attn unavailable:", type(e).__name__)
