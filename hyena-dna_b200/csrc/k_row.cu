#include "launch.h"
namespace hy {

template <int MODE>
static cudaError_t go(const PassArgs& a, int rows, cudaStream_t s) {
  const int M1 = 1 << a.logM1;
  const int nwarps = M1 < 8 ? M1 : 8;
  const int ctas = M1 < 8 ? 1 : M1 / 8;
  size_t elems = (size_t)nwarps * kRowPitch;
  if (MODE == ROW_CONV_FWD) elems += (size_t)nwarps * kM2;
  if (MODE == ROW_CONV_BWD) elems += (size_t)nwarps * 2 * kM2;
  const size_t smem = elems * sizeof(float2);
  auto kern = row_pass_kernel<MODE>;
  cudaError_t e = set_smem(kern, smem);
  if (e != cudaSuccess) return e;
  prof_begin(K_ROW + MODE, s);
  kern<<<dim3(ctas, rows), nwarps * 32, smem, s>>>(a);
  prof_end(K_ROW + MODE, s);
  return cudaGetLastError();
}

cudaError_t launch_row_pass(int mode, const PassArgs& a, int rows, cudaStream_t s) {
  switch (mode) {
    case ROW_FILTER: return go<ROW_FILTER>(a, rows, s);
    case ROW_CONV_FWD: return go<ROW_CONV_FWD>(a, rows, s);
    case ROW_CONV_BWD: return go<ROW_CONV_BWD>(a, rows, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace hy
