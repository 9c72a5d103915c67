// Implicit Hyena filter: positional features -> Sin-MLP -> exponential modulation.
//
// Reference semantics (src/models/sequence/hyena.py):
//   :109-131 PositionalEmbedding  z (L,E), t (L,)           -- read as tensors, never regenerated
//   :96-106  Sin                  sin(freq * x), ONE freq vector shared by the three activations
//   :199-215 implicit_filter      Linear(E,N) Sin Linear(N,N) Sin Linear(N,N) Sin Linear(N,D,no bias)
//   :134-155 ExponentialModulation h * (exp(-t * |deltas|) + shift)
//   :229-238 HyenaFilter.filter
// Output layout is channel-major k[c][t] (D, L) so the FFT column pass reads it coalesced.
//
// fp32 CUDA-core version (N = filter_order = 64).  The MLP is ~50 kflop per position; it is
// evaluated once per step, not per batch element.
#pragma once
#include "common.cuh"

namespace hy {

constexpr int kFN = 64;          // filter_order supported by these kernels
constexpr int kFwdTP = 64;       // positions per CTA, forward
constexpr int kBwdTP = 32;       // positions per tile, backward
constexpr int kMaxE = 16;        // emb_dim limit (odd, >= 3)

struct FilterParams {
  const float* z;        // (L, E)   rows of pos_emb.z[0, :L]
  const float* t;        // (L,)     pos_emb.t[0, :L, 0]
  const float* W0; const float* b0;   // (N,E), (N)
  const float* W1; const float* b1;   // (N,N), (N)
  const float* W2; const float* b2;   // (N,N), (N)
  const float* W3;                    // (D,N)
  const float* freq;                  // (N)
  const float* deltas;                // (D)
  float shift;
  int modulate;
  int L, E, D;
  int z_stride;          // elements between consecutive positions of z (== E when contiguous)
};

struct FilterGrads {
  float* dW0; float* db0; float* dW1; float* db1; float* dW2; float* db2; float* dW3; float* dfreq;
  float* dz;             // (L, E) or null
  int dz_stride;
};

// dynamic smem (floats): Ws[64*65] + actA[64*(TP+1)] + actB[64*(TP+1)] + zs[TP*E] + fr[64] + bb[64]
__host__ __device__ constexpr size_t filter_fwd_smem(int E) {
  return sizeof(float) * (size_t)(kFN * (kFN + 1) + 2 * kFN * (kFwdTP + 1) + kFwdTP * E + 2 * kFN);
}

__host__ __device__ constexpr size_t filter_bwd_smem(int E) {
  // Ws[64*65] + pre[3] + act[3] + dbuf[2] (each 64*(TP+1)) + dh[256*(TP+1)] + zs[TP*E] + fr[64] + bb[64]
  // + dfreq_s[64] + db_s[3*64] + dW0_s[64*E]
  return sizeof(float) * (size_t)(kFN * (kFN + 1) + 8 * kFN * (kBwdTP + 1) + 256 * (kBwdTP + 1) + kBwdTP * E +
                                  2 * kFN + kFN + 3 * kFN + kFN * E);
}

#ifdef HY_FILTER_KERNEL_TU
// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// hidden layer: out[i][t] = sin(freq[i] * (b[i] + sum_j W[i][j] in[j][t])),  i,t in [0,64)
// optionally also stores the pre-activation.
template <int TP, bool KEEP_PRE>
__device__ __forceinline__ void hidden_layer(const float* __restrict__ Ws /*[N][N+1] smem*/,
                                             const float* __restrict__ bias, const float* __restrict__ freq,
                                             const float* __restrict__ in /*[N][TP+1]*/, float* __restrict__ out,
                                             float* __restrict__ pre) {
  constexpr int TT = TP / 4;                 // threads along t
  constexpr int TI = 256 / TT;               // threads along i
  constexpr int IPT = kFN / TI;              // rows per thread
  const int tt = threadIdx.x % TT, ti = threadIdx.x / TT;
  float acc[IPT][4];
#pragma unroll
  for (int a = 0; a < IPT; ++a)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[a][q] = bias[ti * IPT + a];
  for (int j = 0; j < kFN; ++j) {
    float x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = in[j * (TP + 1) + tt + TT * q];
#pragma unroll
    for (int a = 0; a < IPT; ++a) {
      const float w = Ws[(ti * IPT + a) * (kFN + 1) + j];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[a][q] = fmaf(w, x[q], acc[a][q]);
    }
  }
#pragma unroll
  for (int a = 0; a < IPT; ++a) {
    const int i = ti * IPT + a;
    const float f = freq[i];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (KEEP_PRE) pre[i * (TP + 1) + tt + TT * q] = acc[a][q];
      out[i * (TP + 1) + tt + TT * q] = sinf(f * acc[a][q]);
    }
  }
}

// first layer: out[i][t] = sin(freq[i] * (b0[i] + sum_e W0[i][e] z[t][e]))
template <int TP, bool KEEP_PRE>
__device__ __forceinline__ void first_layer(const FilterParams& P, const float* __restrict__ zs /*[TP][E] smem*/,
                                            const float* __restrict__ freq, float* __restrict__ out,
                                            float* __restrict__ pre) {
  for (int o = threadIdx.x; o < kFN * TP; o += blockDim.x) {
    const int i = o / TP, t = o % TP;
    float acc = __ldg(P.b0 + i);
    for (int e = 0; e < P.E; ++e) acc = fmaf(__ldg(P.W0 + i * P.E + e), zs[t * P.E + e], acc);
    if (KEEP_PRE) pre[i * (TP + 1) + t] = acc;
    out[i * (TP + 1) + t] = sinf(freq[i] * acc);
  }
}

__device__ __forceinline__ void load_weight_smem(float* Ws, const float* __restrict__ W) {
  for (int o = threadIdx.x; o < kFN * kFN; o += blockDim.x) Ws[(o / kFN) * (kFN + 1) + (o % kFN)] = __ldg(W + o);
}


__global__ void __launch_bounds__(256, 2) filter_fwd_kernel(const FilterParams P, float* __restrict__ kout) {
  constexpr int TP = kFwdTP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Ws = reinterpret_cast<float*>(smem_raw);
  float* actA = Ws + kFN * (kFN + 1);
  float* actB = actA + kFN * (TP + 1);
  float* zs = actB + kFN * (TP + 1);
  float* fr = zs + TP * P.E;
  float* bb = fr + kFN;
  const int t0 = blockIdx.x * TP;

  for (int o = threadIdx.x; o < TP * P.E; o += blockDim.x) {
    const int t = o / P.E, e = o % P.E;
    zs[o] = (t0 + t < P.L) ? __ldg(P.z + (size_t)(t0 + t) * P.z_stride + e) : 0.f;
  }
  if (threadIdx.x < kFN) fr[threadIdx.x] = __ldg(P.freq + threadIdx.x);
  __syncthreads();
  first_layer<TP, false>(P, zs, fr, actA, nullptr);
  load_weight_smem(Ws, P.W1);
  if (threadIdx.x < kFN) bb[threadIdx.x] = __ldg(P.b1 + threadIdx.x);
  __syncthreads();
  hidden_layer<TP, false>(Ws, bb, fr, actA, actB, nullptr);
  __syncthreads();
  load_weight_smem(Ws, P.W2);
  if (threadIdx.x < kFN) bb[threadIdx.x] = __ldg(P.b2 + threadIdx.x);
  __syncthreads();
  hidden_layer<TP, false>(Ws, bb, fr, actB, actA, nullptr);
  __syncthreads();

  // final projection to D channels + modulation; thread (tc, tt): 8 channels x 8 positions per 256-channel chunk
  const int tt = threadIdx.x % 8, tc = threadIdx.x / 8;
  float tpos[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int t = t0 + tt + 8 * q;
    tpos[q] = (t < P.L) ? __ldg(P.t + t) : 0.f;
  }
  for (int cb = 0; cb < P.D; cb += 256) {
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[a][q] = 0.f;
    for (int j4 = 0; j4 < kFN / 4; ++j4) {
      float x[4][8];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int q = 0; q < 8; ++q) x[jj][q] = actA[(4 * j4 + jj) * (TP + 1) + tt + 8 * q];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int c = cb + tc + 32 * a;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < P.D) w = __ldg(reinterpret_cast<const float4*>(P.W3 + (size_t)c * kFN) + j4);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          acc[a][q] = fmaf(w.x, x[0][q], fmaf(w.y, x[1][q], fmaf(w.z, x[2][q], fmaf(w.w, x[3][q], acc[a][q]))));
      }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int c = cb + tc + 32 * a;
      if (c >= P.D) continue;
      const float ad = fabsf(__ldg(P.deltas + c));
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int t = t0 + tt + 8 * q;
        if (t < P.L) {
          float h = acc[a][q];
          if (P.modulate) h *= (expf(-tpos[q] * ad) + P.shift);
          kout[(size_t)c * P.L + t] = h;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: dk (D, L) -> grads of W0,b0,W1,b1,W2,b2,W3,freq (and z when requested)
// Persistent CTAs loop over position tiles, keep dW accumulators in registers / shared memory and
// flush once with atomics.  D <= 256 per pass (larger D loops over 256-channel chunks and flushes dW3
// per tile).
// ------------------------------------------------------------------------------------------------

// d_in[j][t] = sum_i W[i][j] * dpre[i][t]   (64 x TP outputs, 8 per thread)
template <int TP>
__device__ __forceinline__ void back_linear(const float* __restrict__ Ws /*[N][N+1]*/, const float* __restrict__ dpre,
                                            float* __restrict__ din) {
  // thread (tj in [0,32): j = 2tj, 2tj+1 ; tt in [0,8): t = tt + 8q, q < TP/8)
  const int tt = threadIdx.x % 8, tj = threadIdx.x / 8;
  float acc[2][TP / 8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int q = 0; q < TP / 8; ++q) acc[a][q] = 0.f;
  for (int i = 0; i < kFN; ++i) {
    const float w0 = Ws[i * (kFN + 1) + 2 * tj], w1 = Ws[i * (kFN + 1) + 2 * tj + 1];
#pragma unroll
    for (int q = 0; q < TP / 8; ++q) {
      const float d = dpre[i * (TP + 1) + tt + 8 * q];
      acc[0][q] = fmaf(w0, d, acc[0][q]);
      acc[1][q] = fmaf(w1, d, acc[1][q]);
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int q = 0; q < TP / 8; ++q) din[(2 * tj + a) * (TP + 1) + tt + 8 * q] = acc[a][q];
}

// dW[i][j] += sum_t dpre[i][t] * ain[j][t]; thread owns i = ti + 16a (a<4), j = tj + 16b (b<4)
template <int TP>
__device__ __forceinline__ void accum_dW(float (&dW)[4][4], const float* __restrict__ dpre,
                                         const float* __restrict__ ain) {
  const int tj = threadIdx.x % 16, ti = threadIdx.x / 16;
  for (int t = 0; t < TP; ++t) {
    float d[4], x[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) d[a] = dpre[(ti + 16 * a) * (TP + 1) + t];
#pragma unroll
    for (int b = 0; b < 4; ++b) x[b] = ain[(tj + 16 * b) * (TP + 1) + t];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) dW[a][b] = fmaf(d[a], x[b], dW[a][b]);
  }
}

// through the activation: dpre = da * f * cos(f*pre); dfreq += da * cos(f*pre) * pre; db += dpre
// (in place on da); rows j handled by thread j (first 64 threads own the per-row reductions)
template <int TP>
__device__ __forceinline__ void back_sin(float* __restrict__ da, const float* __restrict__ pre,
                                         const float* __restrict__ fr, float* __restrict__ dfreq_s,
                                         float* __restrict__ db_s, int tvalid) {
  // 256 threads: row j = tid/4, quarter of t
  const int j = threadIdx.x / 4, part = threadIdx.x % 4;
  const float f = fr[j];
  float sf = 0.f, sb = 0.f;
  for (int t = part; t < TP; t += 4) {
    const float p = pre[j * (TP + 1) + t];
    const float cs = cosf(f * p);
    float d = (t < tvalid) ? da[j * (TP + 1) + t] : 0.f;
    sf = fmaf(d * cs, p, sf);
    d = d * f * cs;
    sb += d;
    da[j * (TP + 1) + t] = d;
  }
  sf += __shfl_xor_sync(0xffffffffu, sf, 1); sf += __shfl_xor_sync(0xffffffffu, sf, 2);
  sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
  if (part == 0) { dfreq_s[j] += sf; db_s[j] += sb; }
}


__global__ void __launch_bounds__(256, 1)
filter_bwd_kernel(const FilterParams P, const float* __restrict__ dk, const FilterGrads G, int ntiles) {
  constexpr int TP = kBwdTP;
  constexpr int AS = kFN * (TP + 1);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Ws = reinterpret_cast<float*>(smem_raw);
  float* pre1 = Ws + kFN * (kFN + 1);
  float* pre2 = pre1 + AS;
  float* pre3 = pre2 + AS;
  float* a1 = pre3 + AS;
  float* a2 = a1 + AS;
  float* a3 = a2 + AS;
  float* dA = a3 + AS;
  float* dB = dA + AS;
  float* dh = dB + AS;                        // [256][TP+1]
  float* zs = dh + 256 * (TP + 1);
  float* fr = zs + TP * P.E;
  float* bb = fr + kFN;
  float* dfreq_s = bb + kFN;
  float* db_s = dfreq_s + kFN;                // [3][64]: layers 0,1,2
  float* dW0_s = db_s + 3 * kFN;              // [64][E]

  for (int o = threadIdx.x; o < kFN; o += blockDim.x) { fr[o] = __ldg(P.freq + o); dfreq_s[o] = 0.f; }
  for (int o = threadIdx.x; o < 3 * kFN; o += blockDim.x) db_s[o] = 0.f;
  for (int o = threadIdx.x; o < kFN * P.E; o += blockDim.x) dW0_s[o] = 0.f;

  float dW1r[4][4], dW2r[4][4], dW3r[8][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { dW1r[a][b] = 0.f; dW2r[a][b] = 0.f; }
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) dW3r[a][b] = 0.f;
  const int nchunks = (P.D + 255) / 256;
  __syncthreads();

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int t0 = tile * TP;
    const int tvalid = min(TP, P.L - t0);
    // ---- recompute the forward activations of this tile
    for (int o = threadIdx.x; o < TP * P.E; o += blockDim.x) {
      const int t = o / P.E, e = o % P.E;
      zs[o] = (t < tvalid) ? __ldg(P.z + (size_t)(t0 + t) * P.z_stride + e) : 0.f;
    }
    __syncthreads();
    first_layer<TP, true>(P, zs, fr, a1, pre1);
    load_weight_smem(Ws, P.W1);
    if (threadIdx.x < kFN) bb[threadIdx.x] = __ldg(P.b1 + threadIdx.x);
    __syncthreads();
    hidden_layer<TP, true>(Ws, bb, fr, a1, a2, pre2);
    __syncthreads();
    load_weight_smem(Ws, P.W2);
    if (threadIdx.x < kFN) bb[threadIdx.x] = __ldg(P.b2 + threadIdx.x);
    __syncthreads();
    hidden_layer<TP, true>(Ws, bb, fr, a2, a3, pre3);
    // da3 accumulator lives in dA: zero it
    for (int o = threadIdx.x; o < AS; o += blockDim.x) dA[o] = 0.f;
    __syncthreads();

    // ---- last layer, 256 channels at a time
    for (int ch = 0; ch < nchunks; ++ch) {
      const int cb = ch * 256;
      // dh[c][t] = dk[c][t] * (exp(-t|delta_c|) + shift)
      for (int o = threadIdx.x; o < 256 * TP; o += blockDim.x) {
        const int cl = o / TP, t = o % TP, c = cb + cl;
        float v = 0.f;
        if (c < P.D && t < tvalid) {
          v = __ldg(dk + (size_t)c * P.L + t0 + t);
          if (P.modulate) v *= (expf(-__ldg(P.t + t0 + t) * fabsf(__ldg(P.deltas + c))) + P.shift);
        }
        dh[cl * (TP + 1) + t] = v;
      }
      __syncthreads();
      {   // dW3[c][j] += sum_t dh[c][t] a3[j][t] ; thread: c = tc + 32a, j = tj + 8b
        const int tj = threadIdx.x % 8, tc = threadIdx.x / 8;
        for (int t = 0; t < TP; ++t) {
          float d[8], x[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) d[a] = dh[(tc + 32 * a) * (TP + 1) + t];
#pragma unroll
          for (int b = 0; b < 8; ++b) x[b] = a3[(tj + 8 * b) * (TP + 1) + t];
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) dW3r[a][b] = fmaf(d[a], x[b], dW3r[a][b]);
        }
        if (nchunks > 1) {     // cannot keep several chunks in registers: flush now
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
              const int c = cb + tc + 32 * a;
              if (c < P.D) atomicAdd(G.dW3 + (size_t)c * kFN + tj + 8 * b, dW3r[a][b]);
              dW3r[a][b] = 0.f;
            }
        }
      }
      {   // da3[j][t] += sum_c W3[c][j] dh[c][t] ; thread: j = 2tj,2tj+1 ; t = tt + 8q
        const int tt = threadIdx.x % 8, tj = threadIdx.x / 8;
        float acc[2][TP / 8];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int q = 0; q < TP / 8; ++q) acc[a][q] = 0.f;
        const int cend = min(256, P.D - cb);
        for (int cl = 0; cl < cend; ++cl) {
          const float2 w = __ldg(reinterpret_cast<const float2*>(P.W3 + (size_t)(cb + cl) * kFN) + tj);
#pragma unroll
          for (int q = 0; q < TP / 8; ++q) {
            const float d = dh[cl * (TP + 1) + tt + 8 * q];
            acc[0][q] = fmaf(w.x, d, acc[0][q]);
            acc[1][q] = fmaf(w.y, d, acc[1][q]);
          }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int q = 0; q < TP / 8; ++q) dA[(2 * tj + a) * (TP + 1) + tt + 8 * q] += acc[a][q];
      }
      __syncthreads();
    }

    // ---- layer 2 (W2: a2 -> pre3)
    back_sin<TP>(dA, pre3, fr, dfreq_s, db_s + 2 * kFN, tvalid);        // dA = dpre3   (Ws holds W2)
    __syncthreads();
    accum_dW<TP>(dW2r, dA, a2);
    back_linear<TP>(Ws, dA, dB);                                        // dB = da2
    __syncthreads();
    // ---- layer 1 (W1: a1 -> pre2)
    back_sin<TP>(dB, pre2, fr, dfreq_s, db_s + 1 * kFN, tvalid);        // dB = dpre2
    load_weight_smem(Ws, P.W1);
    __syncthreads();
    accum_dW<TP>(dW1r, dB, a1);
    back_linear<TP>(Ws, dB, dA);                                        // dA = da1
    __syncthreads();
    // ---- layer 0 (W0: z -> pre1)
    back_sin<TP>(dA, pre1, fr, dfreq_s, db_s, tvalid);                  // dA = dpre1
    __syncthreads();
    for (int o = threadIdx.x; o < kFN * P.E; o += blockDim.x) {
      const int i = o / P.E, e = o % P.E;
      float s = 0.f;
      for (int t = 0; t < tvalid; ++t) s = fmaf(dA[i * (TP + 1) + t], zs[t * P.E + e], s);
      dW0_s[o] += s;
    }
    if (G.dz) {
      for (int o = threadIdx.x; o < tvalid * P.E; o += blockDim.x) {
        const int t = o / P.E, e = o % P.E;
        float s = 0.f;
        for (int i = 0; i < kFN; ++i) s = fmaf(__ldg(P.W0 + i * P.E + e), dA[i * (TP + 1) + t], s);
        G.dz[(size_t)(t0 + t) * G.dz_stride + e] = s;
      }
    }
    __syncthreads();
  }

  // ---- flush
  {
    const int tj = threadIdx.x % 16, ti = threadIdx.x / 16;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        atomicAdd(G.dW1 + (ti + 16 * a) * kFN + tj + 16 * b, dW1r[a][b]);
        atomicAdd(G.dW2 + (ti + 16 * a) * kFN + tj + 16 * b, dW2r[a][b]);
      }
  }
  if (nchunks == 1) {
    const int tj = threadIdx.x % 8, tc = threadIdx.x / 8;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int c = tc + 32 * a;
        if (c < P.D) atomicAdd(G.dW3 + (size_t)c * kFN + tj + 8 * b, dW3r[a][b]);
      }
  }
  for (int o = threadIdx.x; o < kFN; o += blockDim.x) {
    atomicAdd(G.dfreq + o, dfreq_s[o]);
    atomicAdd(G.db0 + o, db_s[o]);
    atomicAdd(G.db1 + o, db_s[kFN + o]);
    atomicAdd(G.db2 + o, db_s[2 * kFN + o]);
  }
  for (int o = threadIdx.x; o < kFN * P.E; o += blockDim.x) atomicAdd(G.dW0 + o, dW0_s[o]);
}

#endif  // HY_FILTER_KERNEL_TU

}  // namespace hy
