#!/usr/bin/env python
"""SM clock and board power each kernel class settles at when it runs ALONE, back to back, for a few seconds (NVML, 5 ms samples) -- which
classes of an act step pull the 1 kW cap.   python tools/power_by_class.py [seconds]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
from bench import ClockSampler  # noqa: E402
from hvla import _native as N, config as C, params as P, synthetic as S  # noqa: E402
from hvla.model import HyperVLA  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
B = 64
M = B * 257
lib = N.lib()
st = int(torch.cuda.current_stream().cuda_stream)
sampler = ClockSampler(0)
sampler.start()


def loop(name, fn, flop=None):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.current_stream().synchronize() if n % 200 == 0 else None
    b.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    w = sampler.window(t0 + 0.5 * secs, t1)            # second half: settled
    us = a.elapsed_time(b) * 1e3 / n
    extra = f", {flop / us / 1e6:7.0f} TFLOP/s" if flop else ""
    print(f"{name:34s} {us:8.1f} us/launch{extra}   SM {w['sm_mhz']:.0f} MHz  {w['power_w']:.0f} W  {w['reasons']}", flush=True)


def gemm(n, k, act):
    A = torch.randn(M, k, device="cuda").to(torch.bfloat16)
    Wt = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    Cc = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
    return lambda: lib.hvla_gemm_bf16(st, A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), Cc.data_ptr(), M, n, k, act)


loop("GEMM q|k|v shape (2304 x 768)", gemm(2304, 768, 0), 2.0 * M * 2304 * 768)
loop("GEMM fc1 shape + GELU (3072 x 768)", gemm(3072, 768, 2), 2.0 * M * 3072 * 768)
loop("GEMM fc2 shape (768 x 3072)", gemm(768, 3072, 0), 2.0 * M * 768 * 3072)
qkv = torch.randn(M, 2304, device="cuda")
qkv[:, :768] *= 0.35
qkv = qkv.to(torch.bfloat16)
out = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)
loop("attention (tcgen05)", lambda: lib.hvla_dino_attention(st, qkv.data_ptr(), out.data_ptr(), B, 1), B * 12 * 4.0 * 257 * 257 * 64)
a8 = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
b8 = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
loop("cuBLAS bf16 8192^3 (the peak's workload)", lambda: torch.matmul(a8, b8), 2.0 * 8192 ** 3)
for nm, n, k in (("q|k|v", 2304, 768), ("fc1", 3072, 768), ("fc2", 768, 3072), ("proj", 768, 768)):      # cuBLAS on the SAME shapes (no bias, no epilogue)
    xa = torch.randn(M, k, device="cuda").to(torch.bfloat16)
    xw = torch.randn(k, n, device="cuda").to(torch.bfloat16)
    xo = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
    loop(f"cuBLAS bf16 {nm} shape ({M} x {n} x {k})", lambda xa=xa, xw=xw, xo=xo: torch.matmul(xa, xw, out=xo), 2.0 * M * n * k)
loop("GEMM proj shape (768 x 768)", gemm(768, 768, 0), 2.0 * M * 768 * 768)
params = P.init_params(2025, "P1")
model = HyperVLA.from_config(C.default_config(), precision="bf16", params=params, device="cuda:0")
inp = S.make_inputs(3, B, B)
bp, tasks, _ = model.create_tasks(instruction_dict=inp["instruction_dict"], initial_state=inp["initial_state"])
rt = model.runtime
img = torch.from_numpy(inp["images"][:, 0] if inp["images"].ndim == 5 else inp["images"]).cuda().contiguous()
emb = rt.dino_forward(img)
loop("DINOv2 forward (64 images)", lambda: rt.dino_forward(img))
loop("base net (64 envs)", lambda: rt.base_act(emb, bp.weights if hasattr(bp, "weights") else bp))
sampler.stop_flag.set()
