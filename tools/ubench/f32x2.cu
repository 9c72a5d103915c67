// Microbenchmark: issue rate of scalar vs packed (f32x2) fp32 math on sm_100a.
// Prints warp-instructions per clock per SM and fp32 lane-ops per clock per SM for each variant.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, unsigned long long* cyc) {
  float2 a[8], b = make_float2(seed, seed * 0.5f), c = make_float2(0.999f, 1.001f);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i, seed - i);
  unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, c.x, b.x); a[i].y = fmaf(a[i].y, c.y, b.y); }           // 2 FFMA
      if (MODE == 1) { a[i] = __ffma2_rn(a[i], c, b); }                                               // 1 FFMA2
      if (MODE == 2) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }                                // 2 FADD
      if (MODE == 3) { a[i] = __fadd2_rn(a[i], b); }                                                  // 1 FADD2
      if (MODE == 4) { a[i].x = a[i].x * c.x; a[i].y = a[i].y * c.y; }                                // 2 FMUL
      if (MODE == 5) { a[i] = __fmul2_rn(a[i], c); }                                                  // 1 FMUL2
      if (MODE == 6) {  // complex multiply by a constant twiddle, scalar: 2 FMUL + 2 FFMA
        float2 v = a[i];
        a[i].x = fmaf(v.x, c.x, -v.y * c.y);
        a[i].y = fmaf(v.x, c.y, v.y * c.x);
      }
      if (MODE == 7) {  // complex multiply, packed: FMUL2 + FFMA2 with a swapped copy
        float2 v = a[i];
        float2 sw = make_float2(v.y, v.x);
        float2 t = __fmul2_rn(sw, make_float2(-c.y, c.y));
        a[i] = __ffma2_rn(v, make_float2(c.x, c.x), t);
      }
      if (MODE == 8) {  // radix-2 butterfly scalar: 4 FADD
        float2 x = a[i], y = a[(i + 4) & 7];
        a[i] = make_float2(x.x + y.x, x.y + y.y);
        a[(i + 4) & 7] = make_float2(x.x - y.x, x.y - y.y);
      }
      if (MODE == 9) {  // radix-2 butterfly packed: FADD2 + FADD2(neg)
        float2 x = a[i], y = a[(i + 4) & 7];
        a[i] = __fadd2_rn(x, y);
        a[(i + 4) & 7] = __fadd2_rn(x, make_float2(-y.x, -y.y));
      }
    }
  }
  unsigned long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_inner, int lane_ops_per_inner, int ctas_per_sm) {
  float* out; unsigned long long* cyc;
  const int grid = 148 * ctas_per_sm;
  cudaMalloc(&out, grid * 256 * sizeof(float)); cudaMalloc(&cyc, 8);
  k<MODE><<<grid, 256>>>(out, 1.0f, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, 1.0f, cyc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  // per SM: ctas_per_sm * 8 warps, each ITERS*8*instr_per_inner instructions, in c cycles
  double wi = (double)ctas_per_sm * 8 * ITERS * 8 * instr_per_inner / (double)c;
  double lo = (double)ctas_per_sm * 8 * 32 * ITERS * 8 * lane_ops_per_inner / (double)c;
  printf("%-28s ctas/SM %d: %8.3f ms  %9llu cyc  warp-instr/clk/SM %.2f  fp32 lane-ops/clk/SM %.1f\n", name, ctas_per_sm, ms, c, wi, lo);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int occ = 1; occ <= 4; occ *= 2) {
    run<0>("2xFFMA", 2, 2, occ);
    run<1>("FFMA2", 1, 2, occ);
    run<2>("2xFADD", 2, 2, occ);
    run<3>("FADD2", 1, 2, occ);
    run<4>("2xFMUL", 2, 2, occ);
    run<5>("FMUL2", 1, 2, occ);
    run<6>("cmul scalar (2FMUL+2FFMA)", 4, 4, occ);
    run<7>("cmul packed (FMUL2+FFMA2+swap)", 2, 4, occ);
    run<8>("butterfly scalar (4 FADD)", 4, 4, occ);
    run<9>("butterfly packed (2 FADD2)", 2, 4, occ);
  }
  return 0;
}
