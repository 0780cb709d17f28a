// Per-step image preprocessing on the GPU: what InferenceWrapper._resize_image does with
// TensorFlow on the host for ONE camera frame (data/utils/hypervla_interface.py:89-121), for B frames at once:
//   tf.image.resize(lanczos3, antialias=True) -> [optional sqrt(0.9) centre crop_and_resize, bilinear] ->
//   round-half-even, clip to [0,255], uint8 -- written straight into the (B,224,224,3) buffer the act step reads.
// The resampling is separable: rows first into a float32 intermediate [B,S,W,3], then columns (the order of TF's
// ScaleAndTranslate gather).  Span starts / normalised weights are computed once per (input size, S) by the caller
// (hvla/preprocess.py) and live on the device.  Every product and sum is rounded separately (__fmul_rn / __fadd_rn,
// taps in span order) so the result is bit-identical to the float32 oracle.  HBM-bound byte work: one thread per
// output element, coalesced along (x, channel).
#pragma once
#include "common.cuh"

namespace hvla {
namespace prep {

__device__ __forceinline__ uint8_t round_clip_u8(float v) {
  return (uint8_t)fminf(fmaxf(rintf(v), 0.f), 255.f);       // tf.round = half to even
}

// rows: tmp[b, oy, xc] = sum_k wy[oy, k] * in[b, sy[oy] + k, xc],  xc over W*3
__global__ void __launch_bounds__(256)
resize_rows_kernel(const uint8_t* __restrict__ in, float* __restrict__ tmp, const int* __restrict__ sy, const float* __restrict__ wy,
                   int span, int H, int WC, int S, int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xc = (int)(idx % WC);
  const int oy = (int)((idx / WC) % S);
  const int64_t b = idx / ((int64_t)WC * S);
  const int r0 = sy[oy];
  const float* w = wy + (int64_t)oy * span;
  const uint8_t* src = in + (b * H + r0) * (int64_t)WC + xc;
  float acc = 0.f;
  for (int k = 0; k < span && r0 + k < H; ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(w + k), (float)__ldg(src + (int64_t)k * WC)));
  tmp[idx] = acc;
}

// columns: out[b, oy, ox, c] = sum_k wx[ox, k] * tmp[b, oy, sx[ox] + k, c]; TO = uint8_t (final) or float (before the crop)
template <typename TO>
__global__ void __launch_bounds__(256)
resize_cols_kernel(const float* __restrict__ tmp, TO* __restrict__ out, const int* __restrict__ sx, const float* __restrict__ wx,
                   int span, int W, int S, int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % 3);
  const int ox = (int)((idx / 3) % S);
  const int64_t row = idx / (3 * (int64_t)S);              // b * S + oy
  const int c0 = sx[ox];
  const float* w = wx + (int64_t)ox * span;
  const float* src = tmp + (row * W + c0) * 3 + c;
  float acc = 0.f;
  for (int k = 0; k < span && c0 + k < W; ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(w + k), __ldg(src + 3 * k)));
  if (sizeof(TO) == 1) out[idx] = (TO)round_clip_u8(acc);
  else out[idx] = (TO)acc;
}

// tf.image.crop_and_resize (bilinear, extrapolation 0) of the SxS float image with one box, then round/clip
__global__ void __launch_bounds__(256)
crop_bilinear_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int S, float y1s, float x1s, float hs, float ws,
                     int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % 3);
  const int ox = (int)((idx / 3) % S);
  const int oy = (int)((idx / (3 * (int64_t)S)) % S);
  const int64_t b = idx / (3 * (int64_t)S * S);
  const float in_y = __fadd_rn(y1s, __fmul_rn((float)oy, hs));
  const float in_x = __fadd_rn(x1s, __fmul_rn((float)ox, ws));
  if (in_y < 0.f || in_y > (float)(S - 1) || in_x < 0.f || in_x > (float)(S - 1)) { out[idx] = 0; return; }
  const int ty = (int)floorf(in_y), by = (int)ceilf(in_y), lx = (int)floorf(in_x), rx = (int)ceilf(in_x);
  const float yl = __fsub_rn(in_y, (float)ty), xl = __fsub_rn(in_x, (float)lx);
  const float* p = img + b * (int64_t)S * S * 3 + c;
  const float tl = p[((int64_t)ty * S + lx) * 3], tr = p[((int64_t)ty * S + rx) * 3];
  const float bl = p[((int64_t)by * S + lx) * 3], br = p[((int64_t)by * S + rx) * 3];
  const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
  const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
  out[idx] = round_clip_u8(__fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl)));
}

}  // namespace prep
}  // namespace hvla
