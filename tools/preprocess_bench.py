"""Times the GPU image preprocessing (hvla_resize_lanczos3) on 64 camera frames of 480x640 and compares with the
CPU oracle on one frame.  Algorithmic bytes: u8 frames in + u8 224x224 frames out (the float32 intermediate is extra)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
from hvla.preprocess import BatchedImagePreprocessor
from oracle import preprocess_oracle as PO
B, H, W = 64, 480, 640
frames = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
for crop in (False, True):
    pre = BatchedImagePreprocessor(224, crop=crop)
    for _ in range(3): pre(frames)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): out = pre(frames)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    alg = B * (H * W * 3 + 224 * 224 * 3)
    real = alg + 2 * B * 224 * W * 3 * 4 + (2 * B * 224 * 224 * 3 * 4 if crop else 0)
    print(f"crop={crop}: {ms*1e3:.1f} us / {B} frames = {B/ms*1e3:.0f} frames/s; algorithmic {alg/ms/1e6:.0f} GB/s, with intermediates {real/ms/1e6:.0f} GB/s")
one = frames[0].cpu().numpy()
t = time.time(); PO.resize_image(one, 224, crop=True); print(f"CPU oracle: {(time.time()-t)*1e3:.1f} ms / frame")
