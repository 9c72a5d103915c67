#include "launch.h"
#include "fused_conv.cuh"
namespace hy {

template <int LOGM1>
static cudaError_t go(const PassArgs& a, int channels, int ch_per_group, cudaStream_t s) {
  using FG = FusedGeo<LOGM1>;
  auto kern = fused_conv_fwd_kernel<LOGM1>;
  cudaError_t e = set_smem(kern, FG::SMEM);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, FG::SMEM);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  e = cudaMemcpyToSymbolAsync(c_fused_args, &a, sizeof(PassArgs), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&channels, (void*)&ch_per_group};
  prof_begin(K_FUSED_FWD, s);
  e = cudaLaunchCooperativeKernel((const void*)kern, dim3(per_sm * sms), dim3(256), args, FG::SMEM, s);
  prof_end(K_FUSED_FWD, s);
  return e != cudaSuccess ? e : cudaGetLastError();
}

template <int LOGM1>
static cudaError_t go_flow(const PassArgs& a, int rows, int dist, int* counters, cudaStream_t s) {
  using FG = FusedGeo<LOGM1>;
  auto kern = flow_conv_fwd_kernel<LOGM1>;
  cudaError_t e = set_smem(kern, FG::SMEM);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, FG::SMEM);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  FlowArgs f{rows, dist, 2 * dist + 2, counters};
  e = cudaMemsetAsync(counters, 0, sizeof(int) * (3 * (size_t)rows + 1), s);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_fused_args, &a, sizeof(PassArgs), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_flow_args, &f, sizeof(FlowArgs), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  prof_begin(K_FUSED_FWD, s);
  // cooperative launch for its co-residency guarantee (the items spin on each other's counters); no grid.sync inside
  void* noargs[1] = {nullptr};                         // the kernel takes no parameters
  e = cudaLaunchCooperativeKernel((const void*)kern, dim3(per_sm * sms), dim3(256), noargs, FG::SMEM, s);
  prof_end(K_FUSED_FWD, s);
  return e != cudaSuccess ? e : cudaGetLastError();
}

// a.A must hold (2*dist+2)*B scratch rows; counters: 3*rows+1 ints of device memory
cudaError_t launch_flow_conv_fwd(const PassArgs& a, int rows, int dist, int* counters, cudaStream_t s) {
  if (a.logM2 != 10 || dist < 1) return cudaErrorInvalidValue;
  switch (a.logM1) {
    case 5: return go_flow<5>(a, rows, dist, counters, s);
    case 6: return go_flow<6>(a, rows, dist, counters, s);
    case 7: return go_flow<7>(a, rows, dist, counters, s);
    case 8: return go_flow<8>(a, rows, dist, counters, s);
    case 9: return go_flow<9>(a, rows, dist, counters, s);
    case 10: return go_flow<10>(a, rows, dist, counters, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_fused_conv_fwd(const PassArgs& a, int channels, int ch_per_group, cudaStream_t s) {
  if (a.logM2 != 10) return cudaErrorInvalidValue;
  switch (a.logM1) {
    case 5: return go<5>(a, channels, ch_per_group, s);
    case 6: return go<6>(a, channels, ch_per_group, s);
    case 7: return go<7>(a, channels, ch_per_group, s);
    case 8: return go<8>(a, channels, ch_per_group, s);
    case 9: return go<9>(a, channels, ch_per_group, s);
    case 10: return go<10>(a, channels, ch_per_group, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace hy
