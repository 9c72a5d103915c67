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

}  // namespace hy
