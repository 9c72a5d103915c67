// In-register radix-R (R = 1,2,4,...,32) complex FFT and the two-stage block FFT built on it.
//
// Every index below is a compile-time constant after template expansion: the 32 complex
// values a thread owns stay in registers, twiddles W_32^j become FFMA immediates, and the
// "bit reversal" of the decimation-in-frequency butterflies is pure register renaming.
#pragma once
#include "common.cuh"

namespace hy {

// cos(2*pi*j/32), j in [0,8]
__host__ __device__ constexpr float qcos32(int j) {
  constexpr float t[9] = {1.0f, 0.9807852804032304f, 0.9238795325112867f, 0.8314696123025452f, 0.7071067811865476f,
                          0.5555702330196022f, 0.3826834323650898f, 0.19509032201612825f, 0.0f};
  return t[j];
}
__host__ __device__ constexpr float cos32(int j) {
  j &= 31;
  return j <= 8 ? qcos32(j) : j <= 16 ? -qcos32(16 - j) : j <= 24 ? -qcos32(j - 16) : qcos32(32 - j);
}
__host__ __device__ constexpr float sin32(int j) { return cos32(j + 24); }

// a * W_32^J (forward, W = exp(-2 pi i/32)) or a * conj(W_32^J) (INV)
template <int J, bool INV>
__device__ __forceinline__ float2 mul_w32(float2 a) {
  constexpr int j = J & 31;
  if constexpr (j == 0) return a;
  else if constexpr (j == 16) return make_float2(-a.x, -a.y);
  else if constexpr (j == 8) return INV ? cmul_i(a) : cmul_negi(a);
  else if constexpr (j == 24) return INV ? cmul_negi(a) : cmul_i(a);
  else {
    constexpr float c = cos32(j);
    constexpr float s = INV ? -sin32(j) : sin32(j);      // multiply by (c - i s): one FMUL2 + one FFMA2
    return cmul_cs(a, c, s);
  }
}

// One decimation-in-frequency pass over the sub-array v[BASE .. BASE+N), recursing down to N = 1.
// On return element X[k] of each length-N0 transform sits at v[BASE + brev(k)].
template <int N, int BASE, bool INV, int TOTAL>
__device__ __forceinline__ void dif(float2 (&v)[TOTAL]) {
  if constexpr (N >= 2) {
    constexpr int H = N / 2;
    static_for<0, H>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      float2 a = v[BASE + i], b = v[BASE + i + H];
      v[BASE + i] = cadd(a, b);
      v[BASE + i + H] = mul_w32<i * (32 / N), INV>(csub(a, b));
    });
    dif<H, BASE, INV, TOTAL>(v);
    dif<H, BASE + H, INV, TOTAL>(v);
  }
}

// Same, but the caller promises v[BASE+H .. BASE+N) == 0 on entry (zero-padded upper half):
// the first butterfly layer degenerates to a copy and one twiddle multiply.
template <int N, int BASE, bool INV, int TOTAL>
__device__ __forceinline__ void dif_upper_zero(float2 (&v)[TOTAL]) {
  if constexpr (N >= 2) {
    constexpr int H = N / 2;
    static_for<0, H>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      v[BASE + i + H] = mul_w32<i * (32 / N), INV>(v[BASE + i]);
    });
    dif<H, BASE, INV, TOTAL>(v);
    dif<H, BASE + H, INV, TOTAL>(v);
  }
}

// ------------------------------------------------------------------------------------------
// Geometry of an N-point FFT spread over N/32 threads (N >= 32), each owning 32 points.
//   stage 1: thread q holds x[R2*n1 + q] in v[n1]; radix-32 over n1; twiddle W_N^{q*k1'}
//   exchange through shared memory (row pitch P = R2+1 complex, conflict free)
//   stage 2: thread q owns the G2 = 32/R2 work items k1' = q + R2*i; radix-R2 over n2
//   result:  X[R2*s + q] ("natural slot s") sits in v[slot<LOGN>(s)]
// For N < 32 a thread owns 32/N whole transforms: transform gi lives in v[gi*N .. gi*N+N) and
// X[k] of transform gi sits in v[gi*N + brev(k)].
template <int LOGN>
struct Geo {
  static constexpr int N = 1 << LOGN;
  static constexpr int R2 = N >= 32 ? N / 32 : 1;
  static constexpr int LOGR2 = ilog2c(R2);
  static constexpr int G2 = 32 / R2;
  static constexpr int P = R2 + 1;                 // exchange row pitch (complex elements)
  static constexpr int THREADS = N >= 32 ? N / 32 : 1;
  __host__ __device__ static constexpr int slot(int s) {
    if (N < 32) return 0;   // not used
    if (R2 == 1) return brev(s, 5);
    return (s % G2) * R2 + brev(s / G2, LOGR2);
  }
  // exchange footprint of one transform, in complex elements (without the per-transform pad)
  __host__ __device__ static constexpr int ex_elems() { return R2 == 1 ? 0 : 32 * P; }
};

// v[idx(s)] *= base * step^s for s in [0,32), building the powers with <= 4 chained roundings.
// idx is a constexpr functor s -> register index.  CONJ multiplies by the conjugates instead.
template <bool CONJ, class IDX>
__device__ __forceinline__ void mul_geometric(float2 (&v)[32], float2 base, float2 s1, float2 s2, float2 s4,
                                              float2 s8, float2 s16, IDX) {
  if (CONJ) { base = cconj(base); s1 = cconj(s1); s2 = cconj(s2); s4 = cconj(s4); s8 = cconj(s8); s16 = cconj(s16); }
  float2 lo[8];
  lo[0] = base;
  lo[1] = cmul(base, s1);
  lo[2] = cmul(base, s2);
  lo[3] = cmul(lo[1], s2);
  lo[4] = cmul(base, s4);
  lo[5] = cmul(lo[1], s4);
  lo[6] = cmul(lo[2], s4);
  lo[7] = cmul(lo[3], s4);
  const float2 s24 = cmul(s8, s16);
  static_for<0, 32>([&](auto s_) {
    constexpr int s = decltype(s_)::value;
    constexpr int r = IDX::at(s);
    float2 w = lo[s & 7];
    if constexpr ((s >> 3) == 1) w = cmul(w, s8);
    if constexpr ((s >> 3) == 2) w = cmul(w, s16);
    if constexpr ((s >> 3) == 3) w = cmul(w, s24);
    v[r] = cmul(v[r], w);
  });
}

template <int LOGN> struct SlotIdx { __host__ __device__ static constexpr int at(int s) { return Geo<LOGN>::slot(s); } };
struct Brev5Idx { __host__ __device__ static constexpr int at(int s) { return brev(s, 5); } };
struct IdentIdx { __host__ __device__ static constexpr int at(int s) { return s; } };

// Two-stage N-point FFT (N = 2^LOGN >= 32) of the 32 values in v (input: natural slot s in v[s]).
//   ex    : this transform's exchange area in shared memory (Geo::ex_elems() complex)
//   q     : this thread's index inside the transform, [0, R2)
//   SYNC  : functor called between the exchange write and read (warp or CTA barrier)
// Output: natural slot s (element R2*s + q) in v[Geo<LOGN>::slot(s)].
template <int LOGN, bool INV, bool UPPER_ZERO, class SYNC>
__device__ __forceinline__ void block_fft(float2 (&v)[32], float2* ex, int q, const float2* __restrict__ tw1024,
                                          SYNC sync) {
  using G = Geo<LOGN>;
  static_assert(LOGN >= 5 && LOGN <= 10, "two-stage FFT covers 32..1024 points");
  if constexpr (UPPER_ZERO) dif_upper_zero<32, 0, INV, 32>(v); else dif<32, 0, INV, 32>(v);
  if constexpr (G::R2 > 1) {
    // twiddle W_N^{q*k1'}: geometric in k1' with ratio W_N^q; exponents taken from the 1024-table
    constexpr int SH = 10 - LOGN;                         // W_N^e = tw1024[e << SH]
    const uint32_t e1 = (uint32_t)q << SH;
    float2 s1 = __ldg(tw1024 + (e1 & 1023u));
    float2 s2 = __ldg(tw1024 + ((2u * e1) & 1023u));
    float2 s4 = __ldg(tw1024 + ((4u * e1) & 1023u));
    float2 s8 = __ldg(tw1024 + ((8u * e1) & 1023u));
    float2 s16 = __ldg(tw1024 + ((16u * e1) & 1023u));
    mul_geometric<INV>(v, make_float2(1.f, 0.f), s1, s2, s4, s8, s16, Brev5Idx{});
    // exchange: ex[k1' * P + q]
    static_for<0, 32>([&](auto k_) {
      constexpr int k1 = decltype(k_)::value;
      ex[k1 * G::P + q] = v[brev(k1, 5)];
    });
    sync();
    static_for<0, G::G2>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      static_for<0, G::R2>([&](auto n_) {
        constexpr int n2 = decltype(n_)::value;
        v[i * G::R2 + n2] = ex[(q + G::R2 * i) * G::P + n2];
      });
    });
    static_for<0, G::G2>([&](auto i_) {
      constexpr int i = decltype(i_)::value;
      dif<G::R2, i * G::R2, INV, 32>(v);
    });
  }
}

// ------------------------------------------------------------------------------------------
// 4096-point FFT spread over 128 threads (32 points each): 4096 = 32 x (32 x 4), two exchanges.
//   in : natural slot s (element 128*s + q) in v[s], q = thread index in [0,128)
//   out: natural slot s in v[Slot4096::at(s)]
//   ex : 32*129 complex of shared memory private to this transform; `sync` is a barrier over the
//        128 threads of the transform.
struct Slot4096 { __host__ __device__ static constexpr int at(int s) { return 4 * (s % 8) + brev(s / 8, 2); } };
constexpr int kEx4096 = 32 * 129;

template <bool INV, class SYNC>
__device__ __forceinline__ void block_fft4096(float2 (&v)[32], float2* ex, int q, const Twiddles& T, SYNC sync) {
  // stage 1: radix-32 over n1 (stride 128); X1[k1'] in v[brev5(k1')]; twiddle W_4096^{q k1'}
  dif<32, 0, INV, 32>(v);
  {
    const uint32_t e = (uint32_t)q;             // exponents modulo 4096, scaled by 2^8 into the 2^20 tables
    float2 s1 = root20(T, (e & 4095u) << 8), s2 = root20(T, ((2u * e) & 4095u) << 8),
           s4 = root20(T, ((4u * e) & 4095u) << 8), s8 = root20(T, ((8u * e) & 4095u) << 8),
           s16 = root20(T, ((16u * e) & 4095u) << 8);
    mul_geometric<INV>(v, make_float2(1.f, 0.f), s1, s2, s4, s8, s16, Brev5Idx{});
  }
  static_for<0, 32>([&](auto k_) {
    constexpr int k1 = decltype(k_)::value;
    ex[k1 * 129 + q] = v[brev(k1, 5)];
  });
  sync();
  // stage 2a: thread (k1' = q & 31, n2'' = q >> 5): radix-32 over n1'' of y[4 n1'' + n2'']; twiddle W_128^{n2'' k1''}
  const int k1p = q & 31, c = q >> 5;
  static_for<0, 32>([&](auto n_) {
    constexpr int n1 = decltype(n_)::value;
    v[n1] = ex[k1p * 129 + 4 * n1 + c];
  });
  sync();                                        // everyone has read before the area is reused
  dif<32, 0, INV, 32>(v);
  {
    const uint32_t e = (uint32_t)c * 8u;        // W_128^c = W_1024^{8c}
    float2 s1 = __ldg(T.tw1024 + (e & 1023u)), s2 = __ldg(T.tw1024 + ((2u * e) & 1023u)),
           s4 = __ldg(T.tw1024 + ((4u * e) & 1023u)), s8 = __ldg(T.tw1024 + ((8u * e) & 1023u)),
           s16 = __ldg(T.tw1024 + ((16u * e) & 1023u));
    mul_geometric<INV>(v, make_float2(1.f, 0.f), s1, s2, s4, s8, s16, Brev5Idx{});
  }
  static_for<0, 32>([&](auto k_) {
    constexpr int k1 = decltype(k_)::value;
    ex[k1p * 129 + 4 * k1 + c] = v[brev(k1, 5)];
  });
  sync();
  // stage 2b: thread q owns (k1', k1'' = c + 4j), j < 8: radix-4 over n2''
  static_for<0, 8>([&](auto j_) {
    constexpr int j = decltype(j_)::value;
    static_for<0, 4>([&](auto n_) {
      constexpr int n = decltype(n_)::value;
      v[4 * j + n] = ex[k1p * 129 + (c + 4 * j) * 4 + n];
    });
  });
  static_for<0, 8>([&](auto j_) { dif<4, 4 * decltype(j_)::value, INV, 32>(v); });
}

}  // namespace hy
