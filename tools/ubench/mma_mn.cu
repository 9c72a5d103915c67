// Probe: tcgen05.mma kind::tf32 with the A operand in tensor memory and the B operand MN-major in shared memory --
// which (LBO, SBO) assignment does the no-swizzle MN-major descriptor want?  One CTA, one K = 8 MMA (plus a 4-step K = 32).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hyena-dna_b200/csrc/tc_prims.cuh"
using namespace hy;

// variant: 0 = MN-major B, desc(lbo = KBLK, sbo = 128); 1 = MN-major B, desc(lbo = 128, sbo = KBLK);
//          2 = K-major B (reference), desc(lbo = 128, sbo = 1024) with K = 32 image
__global__ void __launch_bounds__(160, 1) probe(int variant, int N, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t mbar = tc::smem_u32(&bar);
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) tc::mbar_init(mbar, 1);
  // B[n][k] = (n % 5) - 0.5 * (k % 8) + 0.125 * (k / 8), k < 32
  const uint32_t kblk = (uint32_t)(N / 4) * 128u;
  for (int i = tid; i < N * 32; i += blockDim.x) {
    const int n = i / 32, k = i % 32;
    const float v = (float)(n % 5) - 0.5f * (float)(k % 8) + 0.125f * (float)(k / 8);
    uint32_t off;
    if (variant == 2) off = (n / 8) * 1024 + (k / 4) * 128 + (n % 8) * 16 + (k % 4) * 4;
    else off = (k / 8) * kblk + (n / 4) * 128 + (k % 8) * 16 + (n % 4) * 4;
    *reinterpret_cast<float*>(smem + off) = v;
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_s;
  if (warp < 4) {
    // A[m][k] = (m % 7) + 0.25 * k, k < 32, columns 256..287 of lane m
    uint32_t r[32];
    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint((float)(tid % 7) + 0.25f * (float)k);
    tc::tmem_st32(tmem + ((uint32_t)(32 * warp) << 16) + 256, r);
    tc::tmem_wait_st();
    tc::fence_before_sync();
  }
  __syncthreads();
  if (tid == 128) {
    tc::fence_after_sync();
    const uint32_t sb = tc::smem_u32(smem);
    const uint32_t idesc = (variant == 2) ? tc::make_idesc(N) : tc::make_idesc_major(N, false, true);
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t d;
      if (variant == 0) d = tc::make_desc_ls(sb + ks * kblk, kblk, 128u);
      else if (variant == 1) d = tc::make_desc_ls(sb + ks * kblk, 128u, kblk);
      else d = tc::make_desc_ls(sb + ks * 256, 128u, 1024u);
      tc::mma_tf32_ts(tmem, tmem + 256 + 8 * ks, d, idesc, ks ? 1u : 0u);
    }
    tc::mma_commit(mbar);
  }
  if (warp < 4) {
    tc::mbar_wait_u(mbar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32_nowait(tmem + ((uint32_t)(32 * warp) << 16) + c0, r);
      tc::tmem_wait_ld();
      for (int j = 0; j < 32 && c0 + j < N; ++j) out[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  for (int N : {64, 256}) {
    for (int variant = 0; variant < 3; ++variant) {
      float* out; cudaMalloc(&out, 128 * N * sizeof(float)); cudaMemset(out, 0, 128 * N * sizeof(float));
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      probe<<<1, 160, 65536>>>(variant, N, out);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> h(128 * N);
      cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < 32; ++k) ref += ((m % 7) + 0.25 * k) * ((n % 5) - 0.5 * (k % 8) + 0.125 * (k / 8));
          maxerr = fmax(maxerr, fabs(ref - h[m * N + n])); maxref = fmax(maxref, fabs(ref));
        }
      printf("N=%d variant %d: %s max|err| %.4g (max|ref| %.4g)  D[0][0..3] = %g %g %g %g\n", N, variant, cudaGetErrorString(e),
             maxerr, maxref, h[0], h[1], h[2], h[3]);
      cudaFree(out);
    }
  }
  return 0;
}
