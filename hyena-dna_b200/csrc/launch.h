// Host-side launch entry points implemented in the k_*.cu translation units.
#pragma once
#include "fft_passes.cuh"
#include "filter_mlp.cuh"
#include "short_conv.cuh"

namespace hy {

// kernel classes for the launch counter / per-launch event timing (api.cu)
enum Kind {
  K_COL_FWD = 0,      // + ColMode  (0..3)
  K_COL_INV = 4,      // + InvMode  (4..8)
  K_ROW = 9,          // + RowMode  (9..11)
  K_FILTER_FWD = 12, K_FILTER_BWD = 13, K_SHORT_BWD = 14, K_TWIDDLE = 15, K_FILTER_TC_PREP = 16, K_FILTER_TC_FWD = 17,
  K_FILTER_TC_BWD = 18, K_FILTER_TC_RED = 19, K_FUSED_FWD = 20, K_COUNT = 21
};
void prof_begin(int kind, cudaStream_t s);     // api.cu: records an event when profiling is on
void prof_end(int kind, cudaStream_t s);       // api.cu: records an event when profiling is on; counts the launch
cudaError_t launch_col_fwd(int mode, const PassArgs& a, int rows, cudaStream_t s);
cudaError_t launch_col_inv(int mode, const PassArgs& a, int rows, cudaStream_t s);
cudaError_t launch_row_pass(int mode, const PassArgs& a, int rows, cudaStream_t s);
cudaError_t launch_fused_conv_fwd(const PassArgs& a, int channels, int ch_per_group, cudaStream_t s);   // k_fused.cu
cudaError_t launch_flow_conv_fwd(const PassArgs& a, int rows, int dist, int* counters, cudaStream_t s);     // k_fused.cu
cudaError_t launch_filter_fwd(const FilterParams& P, float* kout, cudaStream_t s);
cudaError_t launch_filter_fwd_tc(const FilterParams& P, float* wimg, float* kout, cudaStream_t s);   // k_filter_tc.cu
size_t filter_tc_wimg_bytes(int D);
cudaError_t launch_filter_bwd_tc(const FilterParams& P, float* wimg, const float* dk, float* dh, float* scratch, cudaStream_t s);
struct RedLaunch { const float* dh; const float* scratch; const float* zT; float* dW0; float* db0; float* dW1; float* db1;
                   float* dW2; float* db2; float* dW3; float* dfreq; int L, D, E; };
cudaError_t launch_filter_red_tc(const RedLaunch& r, cudaStream_t s);   // k_filter_tc.cu
cudaError_t launch_filter_bwd(const FilterParams& P, const float* dk, const FilterGrads& G, cudaStream_t s);
cudaError_t launch_short_bwd(const ShortBwdArgs& a, int B, cudaStream_t s);
cudaError_t launch_twiddle_init(float2* tw1024, float2* twlo, cudaStream_t s);

template <class K>
inline cudaError_t set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return cudaSuccess;
}

}  // namespace hy
