#include "launch.h"
namespace hy {

template <int LOGM1, int LOGM2, int MODE>
static cudaError_t go(const PassArgs& a, int rows, cudaStream_t s) {
  using CG = ColGeo<LOGM1, LOGM2>;
  auto kern = col_inv_kernel<LOGM1, LOGM2, MODE>;
  cudaError_t e = set_smem(kern, CG::SMEM_INV);
  if (e != cudaSuccess) return e;
  prof_begin(K_COL_INV + MODE, s);
  kern<<<dim3(CG::CTAS, rows), CG::THREADS, CG::SMEM_INV, s>>>(a);
  prof_end(K_COL_INV + MODE, s);
  return cudaGetLastError();
}

// M <= 2^16: rows of 1024 (logM1 0..6);  M >= 2^17: rows of 4096 (logM1 5..8)
template <int MODE>
static cudaError_t by_size(const PassArgs& a, int rows, cudaStream_t s) {
  if (a.logM2 == 10) {
    switch (a.logM1) {
      case 0: return go<0, 10, MODE>(a, rows, s);
      case 1: return go<1, 10, MODE>(a, rows, s);
      case 2: return go<2, 10, MODE>(a, rows, s);
      case 3: return go<3, 10, MODE>(a, rows, s);
      case 4: return go<4, 10, MODE>(a, rows, s);
      case 5: return go<5, 10, MODE>(a, rows, s);
      case 6: return go<6, 10, MODE>(a, rows, s);
      case 7: return go<7, 10, MODE>(a, rows, s);
      case 8: return go<8, 10, MODE>(a, rows, s);
      case 9: return go<9, 10, MODE>(a, rows, s);
      case 10: return go<10, 10, MODE>(a, rows, s);
    }
  } else if (a.logM2 == 12) {
    switch (a.logM1) {
      case 5: return go<5, 12, MODE>(a, rows, s);
      case 6: return go<6, 12, MODE>(a, rows, s);
      case 7: return go<7, 12, MODE>(a, rows, s);
      case 8: return go<8, 12, MODE>(a, rows, s);
    }
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_col_inv(int mode, const PassArgs& a, int rows, cudaStream_t s) {
  switch (mode) {
    case INV_CONV_FWD: return by_size<INV_CONV_FWD>(a, rows, s);
    case INV_BWD_DG: return by_size<INV_BWD_DG>(a, rows, s);
    case INV_DK: return by_size<INV_DK>(a, rows, s);
    case INV_PLAIN_FWD: return by_size<INV_PLAIN_FWD>(a, rows, s);
    case INV_PLAIN_BWD: return by_size<INV_PLAIN_BWD>(a, rows, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace hy
