// Shared helpers for the sm_100a Hyena long-convolution kernels.
//
// Nothing in this directory includes torch headers: the library is a plain C-ABI
// shared object (include/hyena_b200.h) and PyTorch only hands it device pointers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace hy {

// ---------------------------------------------------------------- complex helpers
// A complex value is one 64-bit register pair and every helper below is written on the sm_100 packed fp32
// instructions (add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2): a complex add is ONE instruction and a
// complex multiply TWO.  Lane swaps and sign flips of an operand ((y, x), (-x, y), ...) are operand modifiers of
// the packed instructions (.F32x2.LO_HI, .NP), so multiplications by +-i and conjugates stay free.
// Rounding is identical to the scalar forms fmaf(a.x, b.x, -(a.y * b.y)) etc.
// HY_SCALAR_COMPLEX (per translation unit) selects the scalar forms instead: measured on B200 (profiles/r2_packed_ab.txt)
// the packed forms help the column passes (issue-slot bound: loads, stores and integer address math share the port with
// the butterflies) and hurt the row passes (fp32-pipe bound: FADD2/FFMA2 issue at half rate, same lane throughput, and
// the register-pair constraints cost MOVs there).
#ifndef HY_SCALAR_COMPLEX
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return __ffma2_rn(make_float2(a.x, a.x), b, __fmul2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x)));
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return __ffma2_rn(make_float2(b.x, b.x), a, __fmul2_rn(make_float2(b.y, b.y), make_float2(a.y, -a.x)));
}
// a * (c - i s) with compile-time friendly real constants c, s
__device__ __forceinline__ float2 cmul_cs(float2 a, float c, float s) {
  return __ffma2_rn(a, make_float2(c, c), __fmul2_rn(make_float2(a.y, a.x), make_float2(s, -s)));
}
// elementwise (not complex) product / fma of two sample pairs
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float2 cmul_cs(float2 a, float c, float s) {
  return make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -a.x * s));
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// -i * a
__device__ __forceinline__ float2 cmul_negi(float2 a) { return make_float2(a.y, -a.x); }
// +i * a
__device__ __forceinline__ float2 cmul_i(float2 a) { return make_float2(-a.y, a.x); }

// ---------------------------------------------------------------- compile-time loop
template <int I> struct IC { static constexpr int value = I; };
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(IC<I>{});
    static_for<I + 1, N>(f);
  }
}

__host__ __device__ constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n >> 1); }
__host__ __device__ constexpr int brev(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

// ---------------------------------------------------------------- twiddle tables (device memory)
// tw1024[j] = exp(-2*pi*i * j / 1024)      j in [0,1024)   (also the "hi" table of the 2^20 roots)
// twlo[j]   = exp(-2*pi*i * j / 2^20)      j in [0,1024)   ("lo" table: W_{2^20}^e = tw1024[e>>10] * twlo[e&1023])
struct Twiddles {
  const float2* tw1024;
  const float2* twlo;
};

// W_{2^20}^{e20}, e20 in [0, 2^20)
__device__ __forceinline__ float2 root20(const Twiddles& T, uint32_t e20) {
  float2 hi = __ldg(T.tw1024 + (e20 >> 10));
  float2 lo = __ldg(T.twlo + (e20 & 1023u));
  return cmul(hi, lo);
}

// ---------------------------------------------------------------- cache-hinted global access
__device__ __forceinline__ float2 ld_stream2(const float2* p) {   // read-once data: do not keep in L1
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---------------------------------------------------------------- cp.async (LDGSTS) helpers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;                      // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace hy
