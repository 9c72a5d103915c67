// Forward long convolution as ONE cooperative kernel per operator call: for each group of rows small enough for the
// inter-pass scratch to stay in L2, pass 1 -> grid.sync -> pass 2 -> grid.sync -> pass 3, each pass executed by the
// persistent CTAs looping over the tiles the stand-alone kernels would have been launched with (same bodies,
// fft_passes.cuh).  With one kernel per pass the scratch (8 MB per row, written and re-read between passes) goes
// through HBM -- 25.8 GB of the 72.6 GB of DRAM traffic per step (profiles/r1_launches_step_summary.txt); here the
// producer and the consumer of a scratch row are a grid barrier apart and the working set of a group is a few tens
// of MB.  Experimental: enabled with HYENA_B200_FUSED=1 (see api.cu).
#pragma once
#include <cooperative_groups.h>

#include "fft_passes.cuh"

namespace hy {

template <int LOGM1>
struct FusedGeo {
  using CG = ColGeo<LOGM1, 10>;
  static constexpr int M1 = 1 << LOGM1;
  static constexpr int TILES = CG::CTAS;               // column tiles per row (passes 1 and 3)
  static constexpr int ROWCTAS = M1 / 8;               // 8-row CTAs per row (pass 2)
  static constexpr size_t ROWSMEM = row_smem_elems<ROW_CONV_FWD, 10>(8) * sizeof(float2);
  static constexpr size_t SMEM = (CG::SMEM_FWD > CG::SMEM_INV ? CG::SMEM_FWD : CG::SMEM_INV) > ROWSMEM
                                     ? (CG::SMEM_FWD > CG::SMEM_INV ? CG::SMEM_FWD : CG::SMEM_INV) : ROWSMEM;
  static_assert(LOGM1 >= 5, "fused kernel covers the thread-group column FFTs (M1 >= 32)");
};

// The pass arguments live in __constant__ memory (set with cudaMemcpyToSymbolAsync on the launch stream) so that the
// three phases can be separate noinline device functions -- each with its own register allocation, exactly like the
// stand-alone kernels -- and still read the arguments as constant-bank operands.  (Inlined into one body ptxas spilled
// ~1.5 KB per thread; passing the struct to noinline functions cost ~30 registers of address/field loads in pass 3.)
// One argument block per device: concurrent fused launches on several streams of one device are not supported.
__constant__ PassArgs c_fused_args;

template <int LOGM1>
__device__ __noinline__ void fused_phase1(int bx, int by, int c0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  col_fwd_body<LOGM1, 10, COL_GATE>(c_fused_args, bx, by, smem_raw, c0);
}
__device__ __noinline__ void fused_phase2(int bx, int by, int c0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  row_pass_body<ROW_CONV_FWD, 10>(c_fused_args, bx, by, smem_raw, c0);
}
template <int LOGM1>
__device__ __noinline__ void fused_phase3(int bx, int by, int c0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  col_inv_body<LOGM1, 10, INV_CONV_FWD>(c_fused_args, bx, by, smem_raw, c0);
}

template <int LOGM1>
__global__ void __launch_bounds__(256, 2) fused_conv_fwd_kernel(const int channels, const int ch_per_group) {
  namespace cgs = cooperative_groups;
  cgs::grid_group grid = cgs::this_grid();
  using FG = FusedGeo<LOGM1>;
  const int B = c_fused_args.B;
  for (int c0 = 0; c0 < channels; c0 += ch_per_group) {
    const int n = (channels - c0 < ch_per_group) ? channels - c0 : ch_per_group;
    const int rows = n * B;
    for (int w = blockIdx.x; w < FG::TILES * rows; w += gridDim.x) {
      fused_phase1<LOGM1>(w % FG::TILES, w / FG::TILES, c0);
      __syncthreads();
    }
    grid.sync();
    for (int w = blockIdx.x; w < FG::ROWCTAS * rows; w += gridDim.x) {
      fused_phase2(w % FG::ROWCTAS, w / FG::ROWCTAS, c0);
      __syncthreads();
    }
    grid.sync();
    for (int w = blockIdx.x; w < FG::TILES * rows; w += gridDim.x) {
      fused_phase3<LOGM1>(w % FG::TILES, w / FG::TILES, c0);
      __syncthreads();
    }
    grid.sync();                                        // the scratch rows are reused by the next group
  }
}


// ------------------------------------------------------------------------------------------------
// Dataflow variant (HYENA_B200_FUSED=2): the same three phase bodies, but no grid-wide barriers.
//
// Work is a single ordered list of tile items, handed out by an atomic ticket.  "Stage" s of the list holds, in this
// order, the pass-3 tiles of row s-2*DIST, the pass-2 row-CTAs of row s-DIST and the pass-1 tiles of row s, so a
// row's consumers sit DIST stages (DIST * (2*TILES+ROWCTAS) items, more than the ~300 items in flight) behind its
// producers and normally never wait.  Correctness does not depend on that distance: every item first waits for the
// per-row completion counter of the phase it consumes,
//     pass 2 of row r : cnt1[r] == TILES     (all column tiles of the row written)
//     pass 3 of row r : cnt2[r] == ROWCTAS   (all k1-row pairs of the row written)
//     pass 1 of row r : cnt3[r-RING] == TILES (the scratch slot r % RING has been drained)
// and an item only ever depends on items with a smaller ticket, which are finished or running on a co-resident CTA
// (cooperative launch), so the smallest unfinished ticket can always make progress: no deadlock.
// The scratch is a ring of RING = 2*DIST+2 rows (8 MB each at M = 2^20): producer and consumer of a row are a few
// hundred items apart and the ring fits L2.
// Publish: bar.sync; thread 0: fence.acq_rel.gpu (cumulative over the CTA's writes) ; red.add.  Wait: thread 0 spins
// with ld.acquire.gpu ; fence ; bar.sync -- the pattern cooperative_groups' grid.sync uses, per row instead of per grid.
// ------------------------------------------------------------------------------------------------
struct FlowArgs {
  int rows;        // B * D logical rows (row = channel * B + batch)
  int dist;        // stages between producer and consumer phases
  int ring;        // scratch rows in flight (a.A holds at least ring rows)
  int* cnt;        // [3 * rows + 1] zeroed before launch: cnt1 | cnt2 | cnt3 | ticket
};
__constant__ FlowArgs c_flow_args;

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void flow_wait(const int* counter, int need) {
  if (threadIdx.x == 0) {
    while (ld_acquire_gpu(counter) < need) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void flow_publish(int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1);
  }
}

template <int LOGM1>
__global__ void __launch_bounds__(256, 2) flow_conv_fwd_kernel() {
  using FG = FusedGeo<LOGM1>;
  constexpr int T = 2 * FG::TILES + FG::ROWCTAS;       // items per stage
  __shared__ int s_ticket;
  const int R = c_flow_args.rows, DIST = c_flow_args.dist, RING = c_flow_args.ring;
  const int B = c_fused_args.B;
  int* cnt1 = c_flow_args.cnt;
  int* cnt2 = cnt1 + R;
  int* cnt3 = cnt2 + R;
  int* ticket = cnt3 + R;
  const int total = (R + 2 * DIST) * T;
  while (true) {
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int item = s_ticket;
    __syncthreads();                                    // s_ticket may be rewritten from here on
    if (item >= total) break;
    const int stage = item / T, j = item - stage * T;
    int phase, row, tile;
    if (j < FG::TILES) { phase = 3; row = stage - 2 * DIST; tile = j; }
    else if (j < FG::TILES + FG::ROWCTAS) { phase = 2; row = stage - DIST; tile = j - FG::TILES; }
    else { phase = 1; row = stage; tile = j - FG::TILES - FG::ROWCTAS; }
    if (row < 0 || row >= R) continue;
    // logical row -> (channel, batch); scratch slot -> the body's row index `by` (see col_fwd_body: ci = by / B,
    // c = c0 + c0x + ci, scratch row = by): by = slot * B + batch, c0x = channel - slot
    const int ch = row / B, b = row - ch * B;
    const int slot = row % RING;
    const int by = slot * B + b, c0x = ch - slot;
    if (phase == 1) {
      if (row >= RING) flow_wait(cnt3 + row - RING, FG::TILES);
      fused_phase1<LOGM1>(tile, by, c0x);
      flow_publish(cnt1 + row);
    } else if (phase == 2) {
      flow_wait(cnt1 + row, FG::TILES);
      fused_phase2(tile, by, c0x);
      flow_publish(cnt2 + row);
    } else {
      flow_wait(cnt2 + row, FG::ROWCTAS);
      fused_phase3<LOGM1>(tile, by, c0x);
      flow_publish(cnt3 + row);
    }
  }
}

}  // namespace hy
