// Batched action post-processing on the GPU: what InferenceWrapper.step does on the host
// after the model call for ONE env (data/utils/hypervla_interface.py:219-300), for B envs at once:
// un-normalise -> temporal action ensembling (data/utils/action_ensemble.py:6-27) -> euler->axis-angle and
// gripper post-processing per policy setup.  One thread per environment; per-env state lives in a device buffer.
#pragma once
#include "common.cuh"

namespace hvla {
namespace post {

constexpr int STATE_FLOATS = 4 * AH * AD + 4;   // history ring [4][4][7] | count | prev_gripper | sticky(on*1000+repeat) | sticky_action

struct PostP {
  const float* raw;        // [B,4,7] model output
  float* state;            // [B, STATE_FLOATS]
  const uint8_t* reset;    // [B] or null: 1 = episode start for this env (clears history and gripper state)
  float* out_raw;          // [B,7]  un-normalised (ensembled) raw action
  float* out_action;       // [B,7]  world_vector(3) | rot_axangle(3) | gripper(1)
  float a[AD], b[AD];      // NORMAL: std, mean ; BOUNDS: p01, p99
  uint8_t mask[AD];
  int B, norm_type;        // 0 NORMAL, 1 BOUNDS
  int ensemble;            // 0: take horizon step 0; 1: ensemble over the last <= 4 predictions
  float temp;
  int policy;              // 0 google_robot, 1 widowx_bridge, 2 libero
  int sticky_repeat;
};

__global__ void postprocess_kernel(PostP p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.B) return;
  float* st = p.state + (int64_t)e * STATE_FLOATS;
  float* hist = st;
  if (p.reset && p.reset[e]) {
    st[4 * AH * AD + 0] = 0.f;                        // count
    st[4 * AH * AD + 1] = __int_as_float(0x7fc00000); // previous gripper action = None
    st[4 * AH * AD + 2] = 0.f;                        // sticky flag / repeat
    st[4 * AH * AD + 3] = 0.f;
  }
  // un-normalise (hypervla_interface.py:219-242)
  float cur[AH][AD];
#pragma unroll
  for (int h = 0; h < AH; ++h)
#pragma unroll
    for (int d = 0; d < AD; ++d) {
      const float r = p.raw[((int64_t)e * AH + h) * AD + d];
      float v = r;
      if (p.mask[d]) {
        if (p.norm_type == 0) v = __fadd_rn(__fmul_rn(r, p.a[d]), p.b[d]);
        else v = __fadd_rn(__fdiv_rn(__fmul_rn(__fadd_rn(r, 1.0f), __fadd_rn(__fsub_rn(p.b[d], p.a[d]), 1e-8f)), 2.0f), p.a[d]);
      }
      cur[h][d] = v;
    }
  float act[AD];
  if (p.ensemble) {
    int count = (int)st[4 * AH * AD];
    const int slot = count & 3;
#pragma unroll
    for (int h = 0; h < AH; ++h)
#pragma unroll
      for (int d = 0; d < AD; ++d) hist[(slot * AH + h) * AD + d] = cur[h][d];
    count += 1;
    st[4 * AH * AD] = (float)count;
    const int num = count < AH ? count : AH;
    // weights[j] = exp(-temp * j) / sum, j = 0 is the OLDEST prediction; it contributes its (num-1-j)-th horizon step
    double wsum = 0.0;
    for (int j = 0; j < num; ++j) wsum += exp(-(double)p.temp * j);
#pragma unroll
    for (int d = 0; d < AD; ++d) {
      double acc = 0.0;
      for (int j = 0; j < num; ++j) {
        const int s = (count - num + j) & 3;
        acc += exp(-(double)p.temp * j) / wsum * (double)hist[(s * AH + (num - 1 - j)) * AD + d];
      }
      act[d] = (float)acc;
    }
  } else {
#pragma unroll
    for (int d = 0; d < AD; ++d) act[d] = cur[0][d];
  }
#pragma unroll
  for (int d = 0; d < AD; ++d) p.out_raw[(int64_t)e * AD + d] = act[d];
  // euler (sxyz) -> axis * angle in float64 (transforms3d.euler.euler2axangle; hypervla_interface.py:263-267)
  double ax[3];
  {
    const double ai = 0.5 * (double)act[3], aj = 0.5 * (double)act[4], ak = 0.5 * (double)act[5];
    const double ci = cos(ai), si = sin(ai), cj = cos(aj), sj = sin(aj), ck = cos(ak), sk = sin(ak);
    const double cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
    double w = cj * cc + sj * ss, x = cj * sc - sj * cs, y = cj * ss + sj * cc, z = cj * cs - sj * sc;
    const double Nq = w * w + x * x + y * y + z * z;
    double theta = 0.0;
    ax[0] = 1.0; ax[1] = 0.0; ax[2] = 0.0;
    if (Nq >= 2.220446049250313e-16) {
      if (Nq != 1.0) { const double s = sqrt(Nq); w /= s; x /= s; y /= s; z /= s; }
      const double len2 = x * x + y * y + z * z;
      if (len2 >= 2.220446049250313e-16 * 2.220446049250313e-16) {
        theta = 2.0 * acos(fmax(fmin(w, 1.0), -1.0));
        const double l = sqrt(len2);
        ax[0] = x / l; ax[1] = y / l; ax[2] = z / l;
      }
    }
    ax[0] *= theta; ax[1] *= theta; ax[2] *= theta;
  }
  float grip;
  if (p.policy == 0) {                                  // google_robot: relative, sticky (hypervla_interface.py:269-293)
    const float current = act[6];
    const float prev = st[4 * AH * AD + 1];
    float rel = (prev != prev) ? 0.f : prev - current;
    st[4 * AH * AD + 1] = current;
    int flag = (int)st[4 * AH * AD + 2];
    int on = flag / 1000, rep = flag % 1000;
    float sticky = st[4 * AH * AD + 3];
    if (fabsf(rel) > 0.5f && !on) { on = 1; sticky = rel; }
    if (on) { rep += 1; rel = sticky; }
    if (rep == p.sticky_repeat) { on = 0; rep = 0; sticky = 0.f; }
    st[4 * AH * AD + 2] = (float)(on * 1000 + rep);
    st[4 * AH * AD + 3] = sticky;
    grip = rel;
  } else if (p.policy == 1) {
    grip = 2.0f * (act[6] > 0.5f ? 1.0f : 0.0f) - 1.0f;   // widowx_bridge: binarise
  } else {
    grip = 2.0f * act[6] - 1.0f;                          // libero
  }
  float* o = p.out_action + (int64_t)e * AD;
  o[0] = act[0]; o[1] = act[1]; o[2] = act[2];
  o[3] = (float)ax[0]; o[4] = (float)ax[1]; o[5] = (float)ax[2];
  o[6] = grip;
}

}  // namespace post
}  // namespace hvla
