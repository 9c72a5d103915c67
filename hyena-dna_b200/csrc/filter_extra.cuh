// Two options of HyenaFilter outside the shipped configs, so that they do not have to raise:
//   * modulation_lr != 0 (src/models/sequence/hyena.py:145-150: `deltas` registered as a Parameter): gradient of
//       k[c][t] = h[c][t] * (exp(-t |delta_c|) + shift)        w.r.t. delta_c
//     d delta_c = sum_t dk[c][t] * h[c][t] * exp(-t |delta_c|) * (-t sign(delta_c)),   h = k / (exp(..) + shift)
//     one CTA per channel, fixed-order block reduction (deterministic);
//   * normalized=True (hyena.py:235-236: `h = h / torch.norm(h, dim=-1, p=1, keepdim=True)` on (1, L, D): an L1
//     normalisation over the CHANNELS of every position): forward k' = k / s, s[t] = sum_c |k[c][t]|; backward
//     dk = (dk' - sign(k') * sum_c dk'[c] k'[c]) / s.  One thread per position, the channel loop strides L (coalesced
//     across the warp), two sweeps over the column.
// Elementwise / reduction work, HBM-bound, off the default path.
#pragma once
#include "common.cuh"

namespace hy {
namespace fx {

__global__ void __launch_bounds__(256) filter_ddelta_kernel(const float* __restrict__ dk, const float* __restrict__ k,
                                                            const float* __restrict__ t, const float* __restrict__ deltas,
                                                            float shift, int L, float* __restrict__ ddelta) {
  const int c = blockIdx.x;
  const float d = __ldg(deltas + c), ad = fabsf(d);
  const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  const float* dkr = dk + (size_t)c * L;
  const float* kr = k + (size_t)c * L;
  float acc = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float tt = __ldg(t + i);
    const float e = expf(-tt * ad), m = e + shift;
    if (m != 0.f) acc = fmaf(dkr[i] * (kr[i] / m), e * (-tt * sg), acc);
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    ddelta[c] = s;
  }
}

__global__ void __launch_bounds__(256) l1norm_fwd_kernel(const float* __restrict__ k, float* __restrict__ out,
                                                         float* __restrict__ norm, int D, int L) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  float s = 0.f;
  for (int c = 0; c < D; ++c) s += fabsf(k[(size_t)c * L + t]);
  norm[t] = s;
  const float inv = 1.0f / s;                              // s == 0 -> inf / nan, exactly as the reference's division
  for (int c = 0; c < D; ++c) out[(size_t)c * L + t] = k[(size_t)c * L + t] * inv;
}

__global__ void __launch_bounds__(256) l1norm_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                         const float* __restrict__ norm, float* __restrict__ dk, int D, int L) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  float dot = 0.f;
  for (int c = 0; c < D; ++c) dot = fmaf(dout[(size_t)c * L + t], out[(size_t)c * L + t], dot);
  const float inv = 1.0f / norm[t];
  for (int c = 0; c < D; ++c) {
    const float y = out[(size_t)c * L + t];
    const float sg = y > 0.f ? 1.f : (y < 0.f ? -1.f : 0.f);
    dk[(size_t)c * L + t] = (dout[(size_t)c * L + t] - sg * dot) * inv;
  }
}

}  // namespace fx
}  // namespace hy
