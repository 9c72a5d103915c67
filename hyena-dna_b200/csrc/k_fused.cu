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
