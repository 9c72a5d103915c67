// Projection GEMMs of the operator on the 5th-generation tensor cores (tcgen05 / TMEM), fp32 accuracy via 3xTF32.
//
// in_proj / out_proj and their input gradients (src/models/sequence/hyena.py:350-351, :391, :440) are all of the form
//     OUT[pos][n] = sum_k ACT[pos][k] * W[n][k]            pos = up to 2^20 sequence positions, K <= 768, N <= 768
// with the activation either row-major (pos, k) (u, dy) or channel-major (k, pos) (y_pre, ds/dp), and the output either
// channel-major (n, pos) (p, dy_pre: what the FFT passes read) or row-major (pos, n) (y, du).  One persistent,
// warp-specialised kernel:
//
//   tile      128 positions (UMMA M = 128, one TMEM lane per position) x NT outputs, K streamed in chunks of 32
//   A operand the activation chunk goes global -> registers -> (hi, lo) tf32 split -> TENSOR MEMORY (tcgen05.st):
//             no shared-memory round trip, and the MMAs read A from TMEM (the shared-memory port is the scarce
//             resource of a tf32 SS-mode MMA: 96 of 128 B/clk at N = 256)
//   B operand the weights, pre-split once per call into hi / lo images in the canonical no-swizzle K-major
//             core-matrix layout (proj_prep_kernel), streamed by TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a
//             ring of shared-memory stages guarded by mbarriers
//   D         fp32 accumulators in TMEM, double buffered so that the epilogue of tile i overlaps the MMAs of tile i+1
//   3xTF32    x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi);  D += Ahi Bhi + Alo Bhi + Ahi Blo   (lo*lo < 2^-22)
//
// Warp roles (320 threads): warps 0-3 convert (thread = position), warps 4-7 epilogue (thread = position),
// warp 8 lane 0 bulk-copy producer, warp 9 lane 0 MMA issuer.
//
// Optional fused prologue (FIR): the activation is ds (B, C, L) and the GEMM consumes dp = transposed 3-tap depthwise
// filter of ds (dp[t] = w2 ds[t] + w1 ds[t+1] + w0 ds[t+2], hyena.py:363-369 backward), so dp never exists in HBM.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"

namespace hy {
namespace pg {

constexpr int kKC = 32;                 // K chunk (one chunk = 4 MMAs of K = 8 per product)
constexpr int kThreads = 320;
constexpr uint32_t kSBO = 1024, kLBO = 128;

// byte offset of element (n, k) inside one (rows x 32) operand image
__host__ __device__ constexpr uint32_t img_off(int n, int k) {
  return (uint32_t)((n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4);
}

enum ActLayout { ACT_ROW = 0 /* (B, L, K): k contiguous */, ACT_CH = 1 /* (B, K, L): position contiguous */ };
enum OutLayout { OUT_CH = 0 /* (B, N, L) */, OUT_ROW = 1 /* (B, L, N) */ };

struct Args {
  const float* act;      // activation, layout per template
  const float* wimg;     // weight images from proj_prep_kernel: [n_tile][k_chunk][hi | lo][NT x 32]
  float* out;
  const float* bias;     // (N) added in the epilogue, or null
  const float* fir;      // (K, 3) taps of the transposed short filter applied to the activation (ACT_CH only), or null
  int B, L, K, N;        // batch, positions per batch (tensor pitch), reduction size, outputs
  int l0, ln;            // positions [l0, l0 + ln) of every batch are processed (host-side chunking of long sequences)
  int kchunks;           // ceil(K / 32)
  int ntiles_n;          // ceil(N / NT)
  int mtiles_per_b;      // ceil(ln / 128)
};

// ------------------------------------------------------------------------------------------------ weight images
// B[n][k] = transposed ? W[k * ldw + n] : W[n * ldw + k];  rows n >= N and columns k >= K are zero
__global__ void proj_prep_kernel(const float* __restrict__ W, int ldw, int transposed, int N, int K, int NT,
                                 float* __restrict__ img) {
  const int kch = (K + kKC - 1) / kKC, ntn = (N + NT - 1) / NT;
  const size_t total = (size_t)ntn * kch * NT * kKC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % kKC);
    const int nn = (int)((i / kKC) % NT);
    const int kc = (int)((i / ((size_t)kKC * NT)) % kch);
    const int nt = (int)(i / ((size_t)kKC * NT * kch));
    const int n = nt * NT + nn, k = kc * kKC + kk;
    float x = 0.f;
    if (n < N && k < K) x = transposed ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k];
    float hi, lo;
    tc::split_tf32(x, hi, lo);
    float* base = img + ((size_t)nt * kch + kc) * 2 * NT * kKC;
    base[img_off(nn, kk) / 4] = hi;
    base[(size_t)NT * kKC + img_off(nn, kk) / 4] = lo;
  }
}

__host__ __device__ constexpr size_t wimg_floats(int N, int K, int NT) {
  return (size_t)((N + NT - 1) / NT) * ((K + kKC - 1) / kKC) * 2 * NT * kKC;
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int NT> struct Cfg {
  static constexpr int STAGES = NT >= 192 ? 4 : 6;
  static constexpr uint32_t STAGE_BYTES = 2u * NT * kKC * 4u;          // hi + lo image of one K chunk
  static constexpr uint32_t D_COLS = NT;                               // per accumulator buffer
  static constexpr uint32_t A_COL0 = 2 * NT;                           // A buffers after the two accumulators
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 256;
  static_assert(2 * NT + 128 <= 512, "two accumulators and two A (hi, lo) chunk buffers must fit tensor memory");
  static_assert(NT % 16 == 0 && NT >= 16 && NT <= 256, "UMMA N");
};

template <int NT, int ACT, int OUT>
__global__ void __launch_bounds__(kThreads, 1) proj_gemm_kernel(const Args a) {
  using C = Cfg<NT>;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)C::STAGES * C::STAGE_BYTES);
  // barrier map: b_full[S] b_empty[S] a_full[2] a_empty[2] d_full[2] d_empty[2]
  uint32_t* tmem_p = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 8);
  const uint32_t sbase = tc::smem_u32(smem);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto B_FULL = [&](int s) { return bar0 + 8u * s; };
  auto B_EMPTY = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  auto A_FULL = [&](int j) { return bar0 + 8u * (2 * C::STAGES + j); };
  auto A_EMPTY = [&](int j) { return bar0 + 8u * (2 * C::STAGES + 2 + j); };
  auto D_FULL = [&](int j) { return bar0 + 8u * (2 * C::STAGES + 4 + j); };
  auto D_EMPTY = [&](int j) { return bar0 + 8u * (2 * C::STAGES + 6 + j); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_p)), "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) { tc::mbar_init(B_FULL(s), 1); tc::mbar_init(B_EMPTY(s), 1); }
    for (int j = 0; j < 2; ++j) {
      tc::mbar_init(A_FULL(j), 128); tc::mbar_init(A_EMPTY(j), 1);
      tc::mbar_init(D_FULL(j), 1); tc::mbar_init(D_EMPTY(j), 128);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_p;

  const int mtiles = a.B * a.mtiles_per_b;
  const long long ntiles = (long long)mtiles * a.ntiles_n;

  if (warp < 4) {
    // ================================================================== converters: thread = position of the tile
    const int row = tid;                                           // 0..127 == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    uint32_t it = 0;                                               // chunk counter over the CTA's whole work list
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int mt = (int)(tile / a.ntiles_n);
      const int b = mt / a.mtiles_per_b, l = a.l0 + (mt - b * a.mtiles_per_b) * 128 + row;
      const bool pv = l < a.l0 + a.ln;
      for (int kc = 0; kc < a.kchunks; ++kc, ++it) {
        float x[kKC];
        const int k0 = kc * kKC;
        if constexpr (ACT == ACT_ROW) {
          const float* src = a.act + ((size_t)b * a.L + l) * a.K + k0;
          if (pv && k0 + kKC <= a.K && (a.K & 3) == 0) {
#pragma unroll
            for (int j = 0; j < kKC / 4; ++j) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j);
              x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < kKC; ++j) x[j] = (pv && k0 + j < a.K) ? __ldg(src + j) : 0.f;
          }
        } else {
          const float* src = a.act + ((size_t)b * a.K + k0) * a.L + l;
          if (a.fir == nullptr) {
#pragma unroll
            for (int j = 0; j < kKC; ++j) x[j] = (pv && k0 + j < a.K) ? __ldg(src + (size_t)j * a.L) : 0.f;
          } else {
            // dp[t] = w2 ds[t] + w1 ds[t+1] + w0 ds[t+2]: the two look-ahead samples come from the next lanes, the
            // last two lanes of a warp read them from memory
#pragma unroll
            for (int j = 0; j < kKC; ++j) {
              const bool kv = k0 + j < a.K;
              const float* s = src + (size_t)j * a.L;
              const float d0 = (pv && kv) ? __ldg(s) : 0.f;
              float d1 = __shfl_down_sync(0xffffffffu, d0, 1), d2 = __shfl_down_sync(0xffffffffu, d0, 2);
              if (lane >= 31) d1 = (kv && l + 1 < a.L) ? __ldg(s + 1) : 0.f;
              if (lane >= 30) d2 = (kv && l + 2 < a.L) ? __ldg(s + 2) : 0.f;
              const float w0 = kv ? __ldg(a.fir + 3 * (k0 + j)) : 0.f, w1 = kv ? __ldg(a.fir + 3 * (k0 + j) + 1) : 0.f,
                          w2 = kv ? __ldg(a.fir + 3 * (k0 + j) + 2) : 0.f;
              x[j] = fmaf(w2, d0, fmaf(w1, d1, w0 * d2));
            }
          }
        }
        uint32_t hi[kKC], lo[kKC];
#pragma unroll
        for (int j = 0; j < kKC; ++j) {
          float h, lw;
          tc::split_tf32(x[j], h, lw);
          hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(lw);
        }
        const int buf = it & 1;
        tc::mbar_wait_u(A_EMPTY(buf), ((it >> 1) & 1) ^ 1);       // MMAs of the previous use of this buffer are done
        tc::fence_after_sync();
        const uint32_t acol = C::A_COL0 + buf * 64;
        tc::tmem_st32(lane_addr + acol, hi);
        tc::tmem_st32(lane_addr + acol + 32, lo);
        tc::tmem_wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(A_FULL(buf));
      }
    }
  } else if (warp < 8) {
    // ================================================================== epilogue: thread = position of the tile
    const int w4 = warp - 4, row = 32 * w4 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * w4) << 16);
    uint32_t tcount = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int mt = (int)(tile / a.ntiles_n), nt = (int)(tile - (long long)mt * a.ntiles_n);
      const int b = mt / a.mtiles_per_b, l = a.l0 + (mt - b * a.mtiles_per_b) * 128 + row;
      const bool pv = l < a.l0 + a.ln;
      const int dbuf = tcount & 1;
      tc::mbar_wait_u(D_FULL(dbuf), (tcount >> 1) & 1);
      tc::fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32_nowait(lane_addr + dbuf * C::D_COLS + c0, r);
        tc::tmem_wait_ld();
        const int n0 = nt * NT + c0;
        if (n0 >= a.N) break;
        if constexpr (OUT == OUT_CH) {
          float* dst = a.out + ((size_t)b * a.N + n0) * a.L + l;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (pv && n0 + j < a.N) {
              float v = __uint_as_float(r[j]);
              if (a.bias) v += __ldg(a.bias + n0 + j);
              dst[(size_t)j * a.L] = v;
            }
          }
        } else {
          float* dst = a.out + ((size_t)b * a.L + l) * a.N + n0;
          if (pv) {
            if (n0 + 32 <= a.N && (a.N & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                       __uint_as_float(r[4 * j + 3]));
                if (a.bias) {
                  const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n0) + j);
                  v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                }
                reinterpret_cast<float4*>(dst)[j] = v;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < a.N) dst[j] = __uint_as_float(r[j]) + (a.bias ? __ldg(a.bias + n0 + j) : 0.f);
            }
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(D_EMPTY(dbuf));
    }
  } else if (warp == 8) {
    // ================================================================== bulk-copy producer (one thread)
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int mt = (int)(tile / a.ntiles_n), nt = (int)(tile - (long long)mt * a.ntiles_n);
        for (int kc = 0; kc < a.kchunks; ++kc, ++it) {
          const int s = it % C::STAGES;
          tc::mbar_wait_u(B_EMPTY(s), ((it / C::STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(B_FULL(s), C::STAGE_BYTES);
          const float* src = a.wimg + ((size_t)nt * a.kchunks + kc) * (C::STAGE_BYTES / 4);
          tc::bulk_g2s(sbase + s * C::STAGE_BYTES, src, C::STAGE_BYTES, B_FULL(s));
        }
      }
    }
  } else {
    // ================================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc(NT);
      uint32_t it = 0, tcount = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        const int dbuf = tcount & 1;
        tc::mbar_wait_u(D_EMPTY(dbuf), ((tcount >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
        tc::fence_after_sync();
        const uint32_t dcol = tmem + dbuf * C::D_COLS;
        for (int kc = 0; kc < a.kchunks; ++kc, ++it) {
          const int s = it % C::STAGES, abuf = it & 1;
          tc::mbar_wait_u(B_FULL(s), (it / C::STAGES) & 1);
          tc::mbar_wait_u(A_FULL(abuf), (it >> 1) & 1);
          tc::fence_after_sync();
          const uint32_t bhi = sbase + s * C::STAGE_BYTES, blo = bhi + C::STAGE_BYTES / 2;
          const uint32_t ahi = tmem + C::A_COL0 + abuf * 64, alo = ahi + 32;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t aa = (pass == 1) ? alo : ahi;
            const uint32_t bb = (pass == 2) ? blo : bhi;
#pragma unroll
            for (int ks = 0; ks < kKC / 8; ++ks)
              tc::mma_tf32_ts(dcol, aa + 8 * ks, tc::make_desc_ls(bb + ks * 2 * kLBO, kLBO, kSBO), idesc,
                              (kc | pass | ks) ? 1u : 0u);
          }
          tc::mma_commit(A_EMPTY(abuf));                           // A chunk buffer free once these MMAs complete
          tc::mma_commit(B_EMPTY(s));                              // and so is the weight stage
        }
        tc::mma_commit(D_FULL(dbuf));
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS) : "memory");
  }
}

}  // namespace pg
}  // namespace hy

namespace hy {
// ================================================================================================ weight gradients
// dW[m][n] = sum_{b,pos} X[b][m][pos] * Y[b][pos][n]: the reduction runs over the (up to 2^20) sequence positions.
//   X channel-major (B, M, L)  -- ds / y_pre: the A operand; thread = row m loads 32 consecutive positions, optional
//                                 transposed short filter (as above), (hi, lo) split, tcgen05.st into tensor memory
//   Y row-major (B, L, N)      -- u / dy: the B operand with N contiguous = MN-major.  Four converter warps split it
//                                 into hi / lo images in the canonical MN-major core-matrix layout (8 positions x 16 bytes
//                                 per core matrix, SBO = 128 B between groups of four n, LBO between blocks of eight
//                                 positions), read by the MMAs through an MN-major shared-memory descriptor
//   split-K   CTA = (m tile of 128 rows, n tile of <= 256 columns, slice of the position chunks); each CTA keeps its
//             accumulator in TMEM for its whole slice and writes one partial (deterministic: summed by wgrad_reduce).
namespace wg {

constexpr int kThreads = 288;             // warps 0-3 A converters + epilogue, 4-7 B converters, 8 MMA issuer
constexpr int kStagesB = 2;
constexpr uint32_t kBStageBytes = 2u * 32u * 256u * 4u;      // hi + lo, 32 positions x 256 columns
constexpr size_t kSmem = (size_t)kStagesB * kBStageBytes + 256;

struct Args {
  const float* X;       // (B, M, L)
  const float* Y;       // (B, L, N)
  const float* fir;     // (M, 3) or null
  float* part;          // (splits, M, N) partial sums
  int B, L, M, N;
  int chunks_per_b;     // ceil(L / 32)
  int mtiles, ntiles, splits;
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStagesB * kBStageBytes);
  // barriers: b_full[2] b_empty[2] a_full[2] a_empty[2] d_full
  uint32_t* tmem_p = reinterpret_cast<uint32_t*>(bars + 9);
  const uint32_t sbase = tc::smem_u32(smem), bar0 = tc::smem_u32(bars);
  auto B_FULL = [&](int s) { return bar0 + 8u * s; };
  auto B_EMPTY = [&](int s) { return bar0 + 8u * (2 + s); };
  auto A_FULL = [&](int j) { return bar0 + 8u * (4 + j); };
  auto A_EMPTY = [&](int j) { return bar0 + 8u * (6 + j); };
  const uint32_t D_FULL = bar0 + 8u * 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_p)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(B_FULL(s), 128); tc::mbar_init(B_EMPTY(s), 1);
      tc::mbar_init(A_FULL(s), 128); tc::mbar_init(A_EMPTY(s), 1);
    }
    tc::mbar_init(D_FULL, 1);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_p;

  // work item of this CTA
  const int split = blockIdx.x % a.splits;
  const int tile = blockIdx.x / a.splits;
  const int mt = tile / a.ntiles, nt = tile - mt * a.ntiles;
  const int n0 = nt * 256;
  const int ncols = min(256, a.N - n0);
  const int nmma = (ncols + 15) & ~15;                                   // UMMA N
  const uint32_t lbo = (uint32_t)(nmma / 4) * 128u;                      // bytes between blocks of eight positions
  const long long total_chunks = (long long)a.B * a.chunks_per_b;
  const long long c_begin = total_chunks * split / a.splits, c_end = total_chunks * (split + 1) / a.splits;

  if (warp < 4) {
    // ---------------------------------------------------------------- A converters: thread = row m of the tile
    const int m = mt * 128 + tid;
    const bool mv = m < a.M;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    if (a.fir && mv) { w0 = __ldg(a.fir + 3 * m); w1 = __ldg(a.fir + 3 * m + 1); w2 = __ldg(a.fir + 3 * m + 2); }
    uint32_t it = 0;
    for (long long c = c_begin; c < c_end; ++c, ++it) {
      const int b = (int)(c / a.chunks_per_b), l0 = (int)(c - (long long)b * a.chunks_per_b) * 32;
      const float* src = a.X + ((size_t)b * a.M + (mv ? m : 0)) * a.L + l0;
      float x[34];
      const bool fast = mv && (l0 + 36 <= a.L) && ((a.L & 3) == 0);
      if (fast) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j);
          x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
        }
        if (a.fir) { const float2 v = __ldg(reinterpret_cast<const float2*>(src + 32)); x[32] = v.x; x[33] = v.y; }
        else { x[32] = 0.f; x[33] = 0.f; }
      } else {
#pragma unroll
        for (int j = 0; j < 34; ++j) x[j] = (mv && l0 + j < a.L && (j < 32 || a.fir)) ? __ldg(src + j) : 0.f;
      }
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float v = a.fir ? fmaf(w2, x[j], fmaf(w1, x[j + 1], w0 * x[j + 2])) : x[j];
        float h, lw;
        tc::split_tf32(v, h, lw);
        hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(lw);
      }
      const int buf = it & 1;
      tc::mbar_wait_u(A_EMPTY(buf), ((it >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t acol = 256 + buf * 64;
      tc::tmem_st32(lane_addr + acol, hi);
      tc::tmem_st32(lane_addr + acol + 32, lo);
      tc::tmem_wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(A_FULL(buf));
    }
    // ---------------------------------------------------------------- epilogue: partial sums of this slice
    if (c_end > c_begin) {
      tc::mbar_wait_u(D_FULL, 0);
      tc::fence_after_sync();
    }
    float* dst = a.part + ((size_t)split * a.M + (mv ? m : 0)) * a.N + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < nmma; c0 += 32) {
      uint32_t r[32];
      if (c_end > c_begin) { tc::tmem_ld32_nowait(lane_addr + c0, r); tc::tmem_wait_ld(); }
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (mv) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < ncols) dst[c0 + j] = __uint_as_float(r[j]);
      }
    }
    tc::fence_before_sync();
  } else if (warp < 8) {
    // ---------------------------------------------------------------- B converters: Y chunk -> MN-major hi / lo images
    const int t = tid - 128;
    const int kq = t & 7, g0 = t >> 3;                                  // position within a block of 8; first n group
    const int ngroups = nmma / 4;
    const bool vec = (a.N & 3) == 0;
    uint32_t it = 0;
    for (long long c = c_begin; c < c_end; ++c, ++it) {
      const int b = (int)(c / a.chunks_per_b), l0 = (int)(c - (long long)b * a.chunks_per_b) * 32;
      const int s = it & 1;
      tc::mbar_wait_u(B_EMPTY(s), ((it >> 1) & 1) ^ 1);
      unsigned char* hi_img = smem + (size_t)s * kBStageBytes;
      unsigned char* lo_img = hi_img + kBStageBytes / 2;
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb) {
        const int l = l0 + kb * 8 + kq;
        const float* src = a.Y + ((size_t)b * a.L + (l < a.L ? l : 0)) * a.N + n0;
        for (int g = g0; g < ngroups; g += 16) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (l < a.L) {
            if (vec && 4 * g + 4 <= ncols) v = __ldg(reinterpret_cast<const float4*>(src) + g);
            else {
              if (4 * g < ncols) v.x = __ldg(src + 4 * g);
              if (4 * g + 1 < ncols) v.y = __ldg(src + 4 * g + 1);
              if (4 * g + 2 < ncols) v.z = __ldg(src + 4 * g + 2);
              if (4 * g + 3 < ncols) v.w = __ldg(src + 4 * g + 3);
            }
          }
          float4 h, lw;
          tc::split_tf32(v.x, h.x, lw.x); tc::split_tf32(v.y, h.y, lw.y);
          tc::split_tf32(v.z, h.z, lw.z); tc::split_tf32(v.w, h.w, lw.w);
          const uint32_t off = (uint32_t)kb * lbo + (uint32_t)g * 128u + (uint32_t)kq * 16u;
          *reinterpret_cast<float4*>(hi_img + off) = h;
          *reinterpret_cast<float4*>(lo_img + off) = lw;
        }
      }
      tc::fence_async_smem();
      tc::mbar_arrive(B_FULL(s));
    }
  } else {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_major(nmma, false, true);
      uint32_t it = 0;
      for (long long c = c_begin; c < c_end; ++c, ++it) {
        const int s = it & 1;
        tc::mbar_wait_u(B_FULL(s), (it >> 1) & 1);
        tc::mbar_wait_u(A_FULL(s), (it >> 1) & 1);
        tc::fence_after_sync();
        const uint32_t bhi = sbase + s * kBStageBytes, blo = bhi + kBStageBytes / 2;
        const uint32_t ahi = tmem + 256 + s * 64, alo = ahi + 32;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t aa = (pass == 1) ? alo : ahi;
          const uint32_t bb = (pass == 2) ? blo : bhi;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc::mma_tf32_ts(tmem, aa + 8 * ks, tc::make_desc_ls(bb + ks * lbo, lbo, 128u), idesc, (it | pass | ks) ? 1u : 0u);
        }
        tc::mma_commit(A_EMPTY(s));
        tc::mma_commit(B_EMPTY(s));
      }
      if (c_end > c_begin) tc::mma_commit(D_FULL);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// dW = sum over splits of part (fixed order: deterministic); transposed: dW is (N, M) and receives part^T
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW, int splits, int M, int N,
                                    int transposed, float beta) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * total + i];
    const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
    float* d = transposed ? dW + (size_t)n * M + m : dW + i;
    *d = (beta != 0.f) ? fmaf(beta, *d, s) : s;
  }
}

}  // namespace wg
}  // namespace hy
