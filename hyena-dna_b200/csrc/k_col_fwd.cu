#include "launch.h"
namespace hy {

template <int LOGM1, int LOGM2, int MODE>
static cudaError_t go(const PassArgs& a, int rows, cudaStream_t s) {
  using CG = ColGeo<LOGM1, LOGM2>;
  auto kern = col_fwd_kernel<LOGM1, LOGM2, MODE>;
  cudaError_t e = set_smem(kern, CG::SMEM_FWD);
  if (e != cudaSuccess) return e;
  prof_begin(K_COL_FWD + MODE, s);
  kern<<<dim3(CG::CTAS, rows), CG::THREADS, CG::SMEM_FWD, s>>>(a);
  prof_end(K_COL_FWD + MODE, s);
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

cudaError_t launch_col_fwd(int mode, const PassArgs& a, int rows, cudaStream_t s) {
  switch (mode) {
    case COL_FILTER: return by_size<COL_FILTER>(a, rows, s);
    case COL_GATE: return by_size<COL_GATE>(a, rows, s);
    case COL_DC: return by_size<COL_DC>(a, rows, s);
    case COL_PLAIN: return by_size<COL_PLAIN>(a, rows, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace hy
