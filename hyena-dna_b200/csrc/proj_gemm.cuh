// Projection GEMMs of the operator on the 5th-generation tensor cores (tcgen05 / TMEM), fp32 accuracy via 3xTF32.
//
// in_proj / out_proj and their input gradients (src/models/sequence/hyena.py:350-351, :391, :440) are all of the form
//     OUT[pos][n] = sum_k ACT[pos][k] * W[n][k]            pos = up to 2^20 sequence positions, K <= 768, N <= 768
// with the activation either row-major (pos, k) (u, dy) or channel-major (k, pos) (y_pre, ds/dp), and the output either
// channel-major (n, pos) (p, dy_pre: what the FFT passes read) or row-major (pos, n) (y, du).  One persistent,
// warp-specialised kernel (proj_gemm_kernel):
//
//   tile      128 positions (UMMA M = 128, one TMEM lane per position) x NT outputs, K streamed in chunks of 32
//   A operand the activation chunk: global -> shared-memory ring by cp.async (four chunks in flight, no registers held
//             across the DRAM latency) -> registers -> (hi, lo) tf32 split -> TENSOR MEMORY (tcgen05.st).  The MMAs
//             read A from TMEM: the shared-memory port is the scarce resource of a tf32 MMA (an SS-mode M128 N256 K8
//             instruction reads 12 KB per 128 cycles), so only B goes through it.
//   B operand the weights, pre-split once per call into hi / lo images in the canonical no-swizzle K-major
//             core-matrix layout (proj_prep_kernel), streamed by TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a
//             ring of shared-memory stages guarded by mbarriers
//   D         fp32 accumulators in TMEM, double buffered: the epilogue of tile i overlaps the MMAs of tile i+1
//   3xTF32    x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi);  D += Ahi Bhi + Alo Bhi + Ahi Blo   (lo*lo < 2^-22)
//
// Warp roles (320 threads): warps 0-3 stage + convert (thread = position), warps 4-7 epilogue (thread = position),
// warp 8 lane 0 bulk-copy producer, warp 9 lane 0 MMA issuer.
//
// Optional fused prologue (FIR): the activation is ds (B, C, L) and the GEMM consumes dp = transposed 3-tap depthwise
// filter of ds (dp[t] = w2 ds[t] + w1 ds[t+1] + w0 ds[t+2], hyena.py:363-369 backward), so dp never exists in HBM.
//
// Weight gradients (wgrad_kernel) at the end of the file.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"

namespace hy {
namespace pg {

constexpr int kKC = 32;                 // K chunk (one chunk = 4 MMAs of K = 8 per product)
constexpr int kThreads = 320;
constexpr int kAStages = 4;             // activation chunks in flight
constexpr uint32_t kSBO = 1024, kLBO = 128;
constexpr uint32_t kAPitchCh = 132 * 4; // ACT_CH staging row: 128 positions + one look-ahead quad (fused FIR)
constexpr uint32_t kAStageBytes = 32 * kAPitchCh;      // 16.5 KB (>= the 16 KB an ACT_ROW tile needs)

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
  int vec;               // 1: 16-byte cp.async staging is legal (alignment / divisibility checked on the host)
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
  static constexpr int STAGES = NT >= 192 ? 2 : 3;                     // weight stages (the weights come from L2)
  static constexpr uint32_t STAGE_BYTES = 2u * NT * kKC * 4u;          // hi + lo image of one K chunk
  static constexpr uint32_t D_COLS = NT;                               // per accumulator buffer
  static constexpr uint32_t A_COL0 = 2 * NT;                           // A buffers after the two accumulators
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr size_t OFF_A = (size_t)STAGES * STAGE_BYTES;        // activation staging ring
  static constexpr size_t OFF_BAR = OFF_A + (size_t)kAStages * kAStageBytes;
  static constexpr size_t OFF_FIR = OFF_BAR + 256;
  static constexpr size_t SMEM = OFF_FIR;                              // + 12 K bytes of taps when the FIR is fused
  static_assert(2 * NT + 128 <= 512, "two accumulators and two A (hi, lo) chunk buffers must fit tensor memory");
  static_assert(NT % 16 == 0 && NT >= 16 && NT <= 256, "UMMA N");
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NT, int ACT, int OUT>
__global__ void __launch_bounds__(kThreads, 1) proj_gemm_kernel(const Args a) {
  using C = Cfg<NT>;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  // barrier map: b_full[S] b_empty[S] a_full[2] a_empty[2] d_full[2] d_empty[2]
  uint32_t* tmem_p = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 8);
  float* fir_s = reinterpret_cast<float*>(smem + C::OFF_FIR);     // (K, 3) taps, FIR only
  const uint32_t sbase = tc::smem_u32(smem);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto B_FULL = [&](int s) { return bar0 + 8u * s; };
  auto B_EMPTY = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  auto A_FULL = [&](int j) { return bar0 + 8u * (2 * C::STAGES + j); };
  auto A_EMPTY = [&](int j) { return bar0 + 8u * (2 * C::STAGES + 2 + j); };
  auto D_FULL = [&](int j) { return bar0 + 8u * (2 * C::STAGES + 4 + j); };
  auto D_EMPTY = [&](int j) { return bar0 + 8u * (2 * C::STAGES + 6 + j); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool use_fir = (ACT == ACT_CH) && a.fir != nullptr;
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_p)), "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (use_fir)
    for (int i = tid; i < 3 * a.K; i += kThreads) fir_s[i] = __ldg(a.fir + i);
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
  const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tiles of this CTA (>= 1)
  const long long nchunks = (long long)my_tiles * a.kchunks;                         // chunks of this CTA

  if (warp < 4) {
    // ================================================================== stage + convert: thread = position of the tile
    const int row = tid;                                           // 0..127 == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    unsigned char* ring = smem + C::OFF_A;

    // chunk index q of this CTA -> (batch, first position of the tile, first k)
    auto locate = [&](long long q, int& b, int& lt, int& k0) {
      const long long tile = blockIdx.x + (q / a.kchunks) * gridDim.x;
      const int mt = (int)(tile / a.ntiles_n);
      b = mt / a.mtiles_per_b;
      lt = a.l0 + (mt - b * a.mtiles_per_b) * 128;
      k0 = (int)(q % a.kchunks) * kKC;
    };
    // fill staging slot q % kAStages with chunk q (asynchronously when the layout allows 16-byte copies)
    auto stage = [&](long long q) {
      int b, lt, k0;
      locate(q, b, lt, k0);
      unsigned char* st = ring + (size_t)(q % kAStages) * kAStageBytes;
      const int lend = a.l0 + a.ln;
      if constexpr (ACT == ACT_ROW) {
        // own row: 8 pieces of 16 B, swizzled by the row so that the read-back (same offsets, 128-byte row pitch) is
        // bank-conflict free; thread-private, so no barrier between fill and use
        const int l = lt + row;
        const float* src = a.act + ((size_t)b * a.L + (l < lend ? l : 0)) * a.K + k0;
        if (a.vec && k0 + kKC <= a.K) {
#pragma unroll
          for (int c = 0; c < 8; ++c) cp_async16(st + row * 128 + ((c ^ (row & 7)) << 4), src + 4 * c, l < lend);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 v;
            v.x = (l < lend && k0 + 4 * c + 0 < a.K) ? __ldg(src + 4 * c + 0) : 0.f;
            v.y = (l < lend && k0 + 4 * c + 1 < a.K) ? __ldg(src + 4 * c + 1) : 0.f;
            v.z = (l < lend && k0 + 4 * c + 2 < a.K) ? __ldg(src + 4 * c + 2) : 0.f;
            v.w = (l < lend && k0 + 4 * c + 3 < a.K) ? __ldg(src + 4 * c + 3) : 0.f;
            *reinterpret_cast<float4*>(st + row * 128 + ((c ^ (row & 7)) << 4)) = v;
          }
        }
      } else {
        // 32 channel rows of 128 (+4 look-ahead) positions; piece p = (channel p / 33, quad p % 33)
        const int npieces = 32 * 33;
        for (int p = tid; p < npieces; p += 128) {
          const int j = p / 33, qd = p - j * 33;
          if (qd == 32 && !use_fir) continue;
          const int l = lt + 4 * qd;
          // in-range test: the 128 tile positions stop at the processed range, the look-ahead quad at the tensor end
          const int lim = (qd == 32) ? a.L : lend;
          const bool kv = k0 + j < a.K;
          const float* src = a.act + ((size_t)b * a.K + (kv ? k0 + j : 0)) * a.L;
          unsigned char* dst = st + j * kAPitchCh + qd * 16;
          if (a.vec) {
            const bool ok = kv && (l + 4 <= lim);
            cp_async16(dst, src + (ok ? l : 0), ok);
          } else {
            float4 v;
            v.x = (kv && l + 0 < lim) ? __ldg(src + l + 0) : 0.f;
            v.y = (kv && l + 1 < lim) ? __ldg(src + l + 1) : 0.f;
            v.z = (kv && l + 2 < lim) ? __ldg(src + l + 2) : 0.f;
            v.w = (kv && l + 3 < lim) ? __ldg(src + l + 3) : 0.f;
            *reinterpret_cast<float4*>(dst) = v;
          }
        }
      }
      cp_async_commit();
    };

    for (int q = 0; q < kAStages - 1; ++q) {
      if (q < nchunks) stage(q); else cp_async_commit();
    }
    for (long long q = 0; q < nchunks; ++q) {
      cp_async_wait_group<kAStages - 2>();                         // chunk q has landed (this thread's pieces)
      if constexpr (ACT == ACT_CH) named_bar_sync(1, 128);         // ... and everybody else's; slot (q-1) % S is free
      if (q + kAStages - 1 < nchunks) stage(q + kAStages - 1); else cp_async_commit();
      const unsigned char* st = ring + (size_t)(q % kAStages) * kAStageBytes;
      float x[kKC];
      if constexpr (ACT == ACT_ROW) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(st + row * 128 + ((c ^ (row & 7)) << 4));
          x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
      } else {
        if (!use_fir) {
#pragma unroll
          for (int j = 0; j < kKC; ++j) x[j] = *reinterpret_cast<const float*>(st + j * kAPitchCh + row * 4);
        } else {
          int b, lt, k0;
          locate(q, b, lt, k0);
          // dp[t] = w2 ds[t] + w1 ds[t+1] + w0 ds[t+2]; ds beyond the tensor end is zero (staged as zero), a position
          // beyond the processed range produces a value nobody stores
#pragma unroll
          for (int j = 0; j < kKC; ++j) {
            const float* s = reinterpret_cast<const float*>(st + j * kAPitchCh) + row;
            const int kk = (k0 + j < a.K) ? k0 + j : 0;
            x[j] = fmaf(fir_s[3 * kk + 2], s[0], fmaf(fir_s[3 * kk + 1], s[1], fir_s[3 * kk] * s[2]));
          }
        }
      }
      uint32_t hi[kKC], lo[kKC];
#pragma unroll
      for (int j = 0; j < kKC; ++j) {
        float hh, lw;
        tc::split_tf32(x[j], hh, lw);
        hi[j] = __float_as_uint(hh); lo[j] = __float_as_uint(lw);
      }
      const uint32_t it = (uint32_t)q;
      const int buf = it & 1;
      tc::mbar_wait_u(A_EMPTY(buf), ((it >> 1) & 1) ^ 1);         // MMAs of the previous use of this buffer are done
      tc::fence_after_sync();
      const uint32_t acol = C::A_COL0 + buf * 64;
      tc::tmem_st32(lane_addr + acol, hi);
      tc::tmem_st32(lane_addr + acol + 32, lo);
      tc::tmem_wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(A_FULL(buf));
    }
    cp_async_wait_all();
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

// ================================================================================================ weight gradients
// dW[m][n] = sum_{b,pos} X[b][m][pos] * Y[b][pos][n]: the reduction runs over the (up to 2^20) sequence positions.
// Computed as its transpose, T[n][m] = sum_pos Y[pos][n] X[m][pos], so that both operands sit in their natural layout
// (a probe on the B200, tools/ubench/mma_mn.cu, showed the MN-major shared-memory descriptor of kind::tf32 returning
// zeros with either LBO/SBO assignment, so no transposed operand is used):
//   A operand = Y^T: 128 columns n of Y per tile = TMEM lanes; thread n reads Y[pos][n] of 32 staged positions
//               (consecutive threads = consecutive n: conflict free), (hi, lo) split, tcgen05.st
//   B operand = X:   128 rows m per tile, K = position contiguous in memory = K-major; four converter warps turn the
//               staged rows into hi / lo K-major core-matrix images (one 16-byte piece = four consecutive positions of
//               one row; optional transposed short filter on the fly from six staged samples)
//   staging   both chunks arrive by cp.async into shared-memory rings, three chunks in flight
//   split-K   CTA = (n tile, m tile, slice of the position chunks); accumulator (128 x 128) in TMEM for the whole
//             slice, one partial per CTA, summed in fixed order by wgrad_reduce_kernel (deterministic, no atomics).
namespace wg {

constexpr int kThreads = 288;             // warps 0-3 Y staging + A conversion + epilogue, 4-7 X staging + B images, 8 MMA
constexpr int kStg = 3;                   // staged chunks in flight
constexpr uint32_t kYPitch = 132 * 4;     // staged Y row: 128 columns (+pad)
constexpr uint32_t kYStage = 32 * kYPitch;            // 32 positions
constexpr uint32_t kXPitch = 36 * 4;      // staged X row: 32 positions + one look-ahead quad (fused FIR)
constexpr uint32_t kXStage = 128 * kXPitch;           // 128 rows
constexpr uint32_t kImg = 128 * 32 * 4;               // one 128 x 32 operand image (16 KB)
constexpr uint32_t kOffY = 0, kOffX = kOffY + kStg * kYStage;
constexpr uint32_t kOffImg = (kOffX + kStg * kXStage + 1023u) & ~1023u;                       // images: [2][hi | lo]
constexpr uint32_t kOffBar = kOffImg + 2 * 2 * kImg;
constexpr size_t kSmem = kOffBar + 256;
static_assert(kOffImg % 1024 == 0, "operand images must start on a core-matrix group boundary");

struct Args {
  const float* X;       // (B, M, L)
  const float* Y;       // (B, L, N)
  const float* fir;     // (M, 3) or null
  float* part;          // (splits, N, M) partial sums of the TRANSPOSED product
  int B, L, M, N;
  int chunks_per_b;     // ceil(L / 32)
  int mtiles, ntiles, splits;
  int vec;              // 16-byte cp.async staging legal
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_p)), "r"(256)
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
  const int nt = tile / a.mtiles, mt = tile - nt * a.mtiles;
  const int n0 = nt * 128, m0 = mt * 128;
  const int nrows = min(128, a.N - n0);                                 // valid accumulator rows (TMEM lanes)
  const int mcols = min(128, a.M - m0);                                 // valid accumulator columns
  const long long total_chunks = (long long)a.B * a.chunks_per_b;
  const long long c_begin = total_chunks * split / a.splits, c_end = total_chunks * (split + 1) / a.splits;
  const long long nchunks = c_end - c_begin;

  if (warp < 4) {
    // ---------------------------------------------------------------- A side: thread = column n of Y = TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    auto stage = [&](long long q) {                                     // Y[l0 .. l0+32)[n0 .. n0+128) -> slot q % kStg
      const long long c = c_begin + q;
      const int b = (int)(c / a.chunks_per_b), l0 = (int)(c - (long long)b * a.chunks_per_b) * 32;
      unsigned char* st = smem + kOffY + (size_t)(q % kStg) * kYStage;
      for (int p = tid; p < 32 * 32; p += 128) {
        const int k = p >> 5, qd = p & 31;
        const int l = l0 + k, n = n0 + 4 * qd;
        const float* src = a.Y + ((size_t)b * a.L + (l < a.L ? l : 0)) * a.N;
        unsigned char* dst = st + k * kYPitch + qd * 16;
        if (a.vec) {
          const bool ok = (l < a.L) && (n + 4 <= a.N);
          cp_async16(dst, src + (ok ? n : 0), ok);
        } else {
          float4 v;
          v.x = (l < a.L && n + 0 < a.N) ? __ldg(src + n + 0) : 0.f;
          v.y = (l < a.L && n + 1 < a.N) ? __ldg(src + n + 1) : 0.f;
          v.z = (l < a.L && n + 2 < a.N) ? __ldg(src + n + 2) : 0.f;
          v.w = (l < a.L && n + 3 < a.N) ? __ldg(src + n + 3) : 0.f;
          *reinterpret_cast<float4*>(dst) = v;
        }
      }
      cp_async_commit();
    };
    for (int q = 0; q < kStg - 1; ++q) {
      if (q < nchunks) stage(q); else cp_async_commit();
    }
    for (long long q = 0; q < nchunks; ++q) {
      cp_async_wait_group<kStg - 2>();
      pg::named_bar_sync(1, 128);
      if (q + kStg - 1 < nchunks) stage(q + kStg - 1); else cp_async_commit();
      const unsigned char* st = smem + kOffY + (size_t)(q % kStg) * kYStage;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float v = *reinterpret_cast<const float*>(st + k * kYPitch + tid * 4);
        float h, lw;
        tc::split_tf32(v, h, lw);
        hi[k] = __float_as_uint(h); lo[k] = __float_as_uint(lw);
      }
      const uint32_t it = (uint32_t)q;
      const int buf = it & 1;
      tc::mbar_wait_u(A_EMPTY(buf), ((it >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t acol = 128 + buf * 64;
      tc::tmem_st32(lane_addr + acol, hi);
      tc::tmem_st32(lane_addr + acol + 32, lo);
      tc::tmem_wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(A_FULL(buf));
    }
    cp_async_wait_all();
    // ---------------------------------------------------------------- epilogue: partial sums of this slice
    if (nchunks > 0) {
      tc::mbar_wait_u(D_FULL, 0);
      tc::fence_after_sync();
    }
    const bool nv = tid < nrows;
    float* dst = a.part + ((size_t)split * a.N + (nv ? n0 + tid : 0)) * a.M + m0;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
      if (nchunks > 0) { tc::tmem_ld32_nowait(lane_addr + c0, r); tc::tmem_wait_ld(); }
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (nv) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < mcols) dst[c0 + j] = __uint_as_float(r[j]);
      }
    }
    tc::fence_before_sync();
  } else if (warp < 8) {
    // ---------------------------------------------------------------- B side: X rows -> K-major hi / lo images
    const int t = tid - 128;
    const bool use_fir = a.fir != nullptr;
    auto stage = [&](long long q) {                                     // X[m0 .. m0+128)[l0 .. l0+32(+4)) -> slot q % kStg
      const long long c = c_begin + q;
      const int b = (int)(c / a.chunks_per_b), l0 = (int)(c - (long long)b * a.chunks_per_b) * 32;
      unsigned char* st = smem + kOffX + (size_t)(q % kStg) * kXStage;
      for (int p = t; p < 128 * 9; p += 128) {
        const int r = p / 9, qd = p - r * 9;
        if (qd == 8 && !use_fir) continue;
        const int m = m0 + r, l = l0 + 4 * qd;
        const float* src = a.X + ((size_t)b * a.M + (m < a.M ? m : 0)) * a.L;
        unsigned char* dst = st + r * kXPitch + qd * 16;
        if (a.vec) {
          const bool ok = (m < a.M) && (l + 4 <= a.L);
          cp_async16(dst, src + (ok ? l : 0), ok);
        } else {
          float4 v;
          v.x = (m < a.M && l + 0 < a.L) ? __ldg(src + l + 0) : 0.f;
          v.y = (m < a.M && l + 1 < a.L) ? __ldg(src + l + 1) : 0.f;
          v.z = (m < a.M && l + 2 < a.L) ? __ldg(src + l + 2) : 0.f;
          v.w = (m < a.M && l + 3 < a.L) ? __ldg(src + l + 3) : 0.f;
          *reinterpret_cast<float4*>(dst) = v;
        }
      }
      cp_async_commit();
    };
    // this thread converts pieces (row r, quad k4) with r % 8 == t % 8: the eight lanes of a quarter warp then write one
    // contiguous 128-byte core matrix (bank-conflict free); 1024 pieces per chunk, 8 per thread
    const int rlo = t & 7, kq = (t >> 3) & 7, rhi0 = t >> 6;            // rows r = rlo + 8 * (rhi0 + 2 i), i < 8
    float w[8][3];
    if (use_fir) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + rlo + 8 * (rhi0 + 2 * i);
#pragma unroll
        for (int j = 0; j < 3; ++j) w[i][j] = (m < a.M) ? __ldg(a.fir + 3 * m + j) : 0.f;
      }
    }
    for (int q = 0; q < kStg - 1; ++q) {
      if (q < nchunks) stage(q); else cp_async_commit();
    }
    for (long long q = 0; q < nchunks; ++q) {
      cp_async_wait_group<kStg - 2>();
      pg::named_bar_sync(2, 128);
      if (q + kStg - 1 < nchunks) stage(q + kStg - 1); else cp_async_commit();
      const unsigned char* st = smem + kOffX + (size_t)(q % kStg) * kXStage;
      const uint32_t it = (uint32_t)q;
      const int s = it & 1;
      tc::mbar_wait_u(B_EMPTY(s), ((it >> 1) & 1) ^ 1);                 // the MMAs that read this image pair are done
      unsigned char* hi_img = smem + kOffImg + (size_t)s * 2 * kImg;
      unsigned char* lo_img = hi_img + kImg;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rlo + 8 * (rhi0 + 2 * i);
        const float* row = reinterpret_cast<const float*>(st + r * kXPitch) + 4 * kq;
        float4 v = *reinterpret_cast<const float4*>(row);
        if (use_fir) {
          const float2 nx = *reinterpret_cast<const float2*>(row + 4);
          const float x4 = nx.x, x5 = nx.y;
          float4 o;
          o.x = fmaf(w[i][2], v.x, fmaf(w[i][1], v.y, w[i][0] * v.z));
          o.y = fmaf(w[i][2], v.y, fmaf(w[i][1], v.z, w[i][0] * v.w));
          o.z = fmaf(w[i][2], v.z, fmaf(w[i][1], v.w, w[i][0] * x4));
          o.w = fmaf(w[i][2], v.w, fmaf(w[i][1], x4, w[i][0] * x5));
          v = o;
        }
        float4 h, lw;
        tc::split_tf32(v.x, h.x, lw.x); tc::split_tf32(v.y, h.y, lw.y);
        tc::split_tf32(v.z, h.z, lw.z); tc::split_tf32(v.w, h.w, lw.w);
        const uint32_t off = pg::img_off(r, 4 * kq);
        *reinterpret_cast<float4*>(hi_img + off) = h;
        *reinterpret_cast<float4*>(lo_img + off) = lw;
      }
      tc::fence_async_smem();
      tc::mbar_arrive(B_FULL(s));
    }
    cp_async_wait_all();
  } else {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc(128);
      for (long long q = 0; q < nchunks; ++q) {
        const uint32_t it = (uint32_t)q;
        const int s = it & 1;
        tc::mbar_wait_u(B_FULL(s), (it >> 1) & 1);
        tc::mbar_wait_u(A_FULL(s), (it >> 1) & 1);
        tc::fence_after_sync();
        const uint32_t bhi = sbase + kOffImg + s * 2 * kImg, blo = bhi + kImg;
        const uint32_t ahi = tmem + 128 + s * 64, alo = ahi + 32;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t aa = (pass == 1) ? alo : ahi;
          const uint32_t bb = (pass == 2) ? blo : bhi;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc::mma_tf32_ts(tmem, aa + 8 * ks, tc::make_desc_ls(bb + ks * 2 * pg::kLBO, pg::kLBO, pg::kSBO), idesc,
                            (it | pass | ks) ? 1u : 0u);
        }
        tc::mma_commit(A_EMPTY(s));
        tc::mma_commit(B_EMPTY(s));
      }
      if (nchunks > 0) tc::mma_commit(D_FULL);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

// dW = sum over splits of part^T (fixed order: deterministic).  part is (splits, N, M); dW is (M, N), or (N, M) when
// `transposed` (then no transposition is left to do)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW, int splits, int M, int N,
                                    int transposed, float beta) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // i indexes the OUTPUT; fetch part[n][m]
    size_t src;
    if (transposed) src = i;                                            // dW (N, M) == part layout
    else { const int m = (int)(i / N), n = (int)(i - (size_t)m * N); src = (size_t)n * M + m; }
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * total + src];
    dW[i] = (beta != 0.f) ? fmaf(beta, dW[i], s) : s;
  }
}

}  // namespace wg
}  // namespace hy
