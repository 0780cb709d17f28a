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

// byte i of a 32-bit word as an exact float without the quarter-rate I2F: 0x4B0000xx is 8388608 + xx
template <int I>
__device__ __forceinline__ float byte_to_float(uint32_t u) {
  return __fsub_rn(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7650 + I)), 8388608.0f);
}

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

// Both passes in one kernel: a CTA owns one output row of one frame; the row pass lands in shared memory (W*3 floats),
// the column pass reads it from there, so the float32 intermediate never touches HBM.  Same arithmetic, same order.
// VEC: W*3 % 4 == 0 (and a 4-byte aligned frame base): the row pass reads four bytes per load.
template <typename TO, bool VEC>
__global__ void __launch_bounds__(256)
resize_fused_kernel(const uint8_t* __restrict__ in, TO* __restrict__ out, const int* __restrict__ sy, const float* __restrict__ wy,
                    int span_y, const int* __restrict__ sx, const float* __restrict__ wx, int span_x, int H, int W, int S) {
  extern __shared__ float row[];                       // [W*3]
  const int oy = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int WC = W * 3;
  const int r0 = sy[oy];
  const float* w = wy + (int64_t)oy * span_y;
  const int taps = min(span_y, H - r0);
  const uint8_t* src = in + (b * H + r0) * (int64_t)WC;
  if (VEC) {
    for (int x4 = threadIdx.x; x4 < WC / 4; x4 += blockDim.x) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
      for (int k = 0; k < taps; ++k) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(src + (int64_t)k * WC) + x4);
        const float wk = __ldg(w + k);
        a0 = __fadd_rn(a0, __fmul_rn(wk, byte_to_float<0>(u)));
        a1 = __fadd_rn(a1, __fmul_rn(wk, byte_to_float<1>(u)));
        a2 = __fadd_rn(a2, __fmul_rn(wk, byte_to_float<2>(u)));
        a3 = __fadd_rn(a3, __fmul_rn(wk, byte_to_float<3>(u)));
      }
      *reinterpret_cast<float4*>(row + 4 * x4) = make_float4(a0, a1, a2, a3);
    }
  } else {
    for (int xc = threadIdx.x; xc < WC; xc += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < taps; ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(w + k), (float)__ldg(src + (int64_t)k * WC + xc)));
      row[xc] = acc;
    }
  }
  __syncthreads();
  TO* dst = out + (b * S + oy) * (int64_t)S * 3;
  for (int o = threadIdx.x; o < S * 3; o += blockDim.x) {
    const int ox = o / 3, c = o - ox * 3;
    const int c0 = sx[ox];
    const float* wc = wx + (int64_t)ox * span_x;
    const int tx = min(span_x, W - c0);
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < tx; ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(wc + k), row[(c0 + k) * 3 + c]));
    if (sizeof(TO) == 1) dst[o] = (TO)round_clip_u8(acc);
    else dst[o] = (TO)acc;
  }
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
