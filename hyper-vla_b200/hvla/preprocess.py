"""Batched, GPU-resident image preprocessing.

Mirrors ``InferenceWrapper._resize_image`` (data/utils/hypervla_interface.py:89-121) for B camera frames at once:
``tf.image.resize(method="lanczos3", antialias=True)`` to ``image_size`` x ``image_size``, optionally the sqrt(0.9)
centre ``tf.image.crop_and_resize`` (bilinear), then round / clip / uint8.  The output is the (B,224,224,3) uint8
CUDA tensor ``HyperVLA.sample_actions`` takes, so the frames never return to the host.

The resampling weights (TensorFlow's ScaleAndTranslate span rule) depend only on the input size; they are computed
here once per size in float32 and cached on the device.  ``padded_resize`` (resize_with_pad, :90-95) is not supported.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_F = np.float32


def _lanczos3(x):
    x = np.abs(x.astype(_F))
    pi = _F(3.14159265359)
    safe = np.where(x <= _F(1e-3), _F(1), x)
    v = _F(3) * np.sin(pi * safe).astype(_F) * np.sin(pi * safe / _F(3)).astype(_F) / (pi * pi * safe * safe)
    v = np.where(x <= _F(1e-3), _F(1), v)
    return np.where(x > _F(3), _F(0), v).astype(_F)


def lanczos3_spans(out_size: int, in_size: int):
    """Span start and normalised weights per output index (antialiased: the kernel widens by in/out when shrinking)."""
    scale = _F(out_size) / _F(in_size)
    inv_scale = _F(1) / scale
    ks = max(inv_scale, _F(1))
    span = min(2 * int(np.ceil(_F(3) * ks)) + 1, in_size)
    starts = np.zeros(out_size, np.int32)
    weights = np.zeros((out_size, span), _F)
    for x in range(out_size):
        sample = _F(_F(x) + _F(0.5)) * inv_scale
        if sample < 0 or sample > in_size:
            continue
        s = min(max(int(np.ceil(sample - _F(3) * ks - _F(0.5))), 0), in_size - 1)
        e = min(max(int(np.floor(sample + _F(3) * ks - _F(0.5))), 0), in_size - 1) + 1
        w = _lanczos3(np.abs((np.arange(s, e).astype(_F) + _F(0.5) - sample) * (_F(1) / ks)))
        tot = _F(0)
        for v in w:
            tot = _F(tot + v)
        if abs(tot) >= _F(1000) * np.finfo(_F).tiny:
            weights[x, :len(w)] = (w * (_F(1) / tot)).astype(_F)
        starts[x] = s
    return starts, weights


class BatchedImagePreprocessor:
    def __init__(self, image_size: int = 224, crop: bool = False, padded_resize: bool = False, device=None):
        import torch
        if padded_resize:
            raise ValueError("padded_resize is not supported by the GPU preprocessing path")
        if not torch.cuda.is_available():
            raise N.HvlaError("no CUDA device: hvla image preprocessing is CUDA-only")
        self.lib = N.lib()
        self.S, self.crop = int(image_size), bool(crop)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._tables = {}
        self._ws = None
        s = np.sqrt(0.9)                                            # hypervla_interface.py:106
        o = (1 - s) / 2
        y1, y2 = _F(o), _F(o + s)
        n = _F(self.S - 1)
        hs = (y2 - y1) * n / n
        self._crop_params = (C.c_float * 4)(float(y1 * n), float(y1 * n), float(hs), float(hs))

    def _table(self, in_size):
        import torch
        if in_size not in self._tables:
            st, w = lanczos3_spans(self.S, in_size)
            self._tables[in_size] = (torch.from_numpy(st).to(self.device), torch.from_numpy(w).to(self.device), w.shape[1])
        return self._tables[in_size]

    def __call__(self, images):
        """images: (B,H,W,3) uint8, numpy or CUDA tensor -> (B,S,S,3) uint8 CUDA tensor (asynchronous)."""
        import torch
        x = images if torch.is_tensor(images) else torch.from_numpy(np.ascontiguousarray(images))
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError("images must be (B,H,W,3) uint8")
        x = x.to(self.device, non_blocking=True).contiguous()
        B, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        out = torch.empty((B, self.S, self.S, 3), dtype=torch.uint8, device=self.device)
        if B == 0:
            return out
        sy, wy, ny = self._table(H)
        sx, wx, nx = self._table(W)
        need = int(self.lib.hvla_resize_workspace_bytes(B, H, W, self.S, int(self.crop)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):       # launches go to the current device's context
            st = self.lib.hvla_resize_lanczos3(
                int(torch.cuda.current_stream(self.device).cuda_stream), x.data_ptr(), B, H, W, self.S, sy.data_ptr(), wy.data_ptr(), ny,
                sx.data_ptr(), wx.data_ptr(), nx, int(self.crop), C.cast(self._crop_params, C.c_void_p), out.data_ptr(),
                self._ws.data_ptr(), need)
        N.check(st, "hvla_resize_lanczos3")
        return out
