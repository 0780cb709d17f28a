#!/usr/bin/env python
"""Times the tcgen05 GEMM (hvla_gemm_bf16) on the DINOv2 shapes of one batch-64 step (CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
from hvla import _native as N  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = B * 257
SHAPES = [("qkv", M, 2304, 768, 0), ("proj", M, 768, 768, 0), ("fc1+gelu", M, 3072, 768, 2), ("fc2", M, 768, 3072, 0)]
lib = N.lib()
st = int(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for name, m, n, k, act in SHAPES:
    A = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    Wt = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    C = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        N.check(lib.hvla_gemm_bf16(st, A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), C.data_ptr(), m, n, k, act), "gemm")
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.hvla_gemm_bf16(st, A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), C.data_ptr(), m, n, k, act)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    med = ts[len(ts) // 2]
    ref = A.float() @ Wt.float().t() + bias
    if act == 2:
        ref = torch.nn.functional.gelu(ref)
    err = (C.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"{name:9s} M={m} N={n} K={k}: {med * 1e3:8.1f} us  {2.0 * m * n * k / med / 1e9:7.1f} TFLOP/s  (min {ts[0] * 1e3:.1f} us)  err {err:.2e}")
