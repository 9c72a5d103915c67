#include "launch.h"
namespace hy {

template <int MODE, int LOGM2>
static cudaError_t go(const PassArgs& a, int rows, cudaStream_t s) {
  using RG = RowGeo<LOGM2>;
  const int M1 = 1 << a.logM1;
  const int rows_cta = (MODE == ROW_CONV_BWD1) ? (RG::ROWS >= 4 ? RG::ROWS / 2 : RG::ROWS) : RG::ROWS;   // 128-thread CTAs
  const int nslots = M1 < rows_cta ? M1 : rows_cta;
  const int ctas = M1 < rows_cta ? 1 : M1 / rows_cta;
  const size_t smem = row_smem_elems<MODE, LOGM2>(nslots) * sizeof(float2);
  auto kern = row_pass_kernel<MODE, LOGM2>;
  cudaError_t e = set_smem(kern, smem);
  if (e != cudaSuccess) return e;
  prof_begin(K_ROW + (MODE == ROW_CONV_BWD1 ? (int)ROW_CONV_BWD : MODE), s);
  kern<<<dim3(ctas, rows), nslots * RG::TPR, smem, s>>>(a);
  prof_end(K_ROW + (MODE == ROW_CONV_BWD1 ? (int)ROW_CONV_BWD : MODE), s);
  return cudaGetLastError();
}

template <int MODE>
static cudaError_t by_len(const PassArgs& a, int rows, cudaStream_t s) {
  if (a.logM2 == 10) return go<MODE, 10>(a, rows, s);
  if (a.logM2 == 12) return go<MODE, 12>(a, rows, s);
  return cudaErrorInvalidValue;
}

cudaError_t launch_row_pass(int mode, const PassArgs& a, int rows, cudaStream_t s) {
  switch (mode) {
    case ROW_FILTER: return by_len<ROW_FILTER>(a, rows, s);
    case ROW_CONV_FWD: return by_len<ROW_CONV_FWD>(a, rows, s);
    case ROW_CONV_BWD: return by_len<ROW_CONV_BWD>(a, rows, s);
    case ROW_CONV_BWD1: return by_len<ROW_CONV_BWD1>(a, rows, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace hy
