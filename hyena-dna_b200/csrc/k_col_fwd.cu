// Pass-1 dispatch over the input mode; the kernels live in k_col_fwd_m*.cu (one translation unit per mode).
#include "launch.h"
namespace hy {
cudaError_t launch_col_fwd(int mode, const PassArgs& a, int rows, cudaStream_t s) {
  switch (mode) {
    case COL_FILTER: return launch_col_fwd_mode<COL_FILTER>(a, rows, s);
    case COL_GATE: return launch_col_fwd_mode<COL_GATE>(a, rows, s);
    case COL_DC: return launch_col_fwd_mode<COL_DC>(a, rows, s);
    case COL_PLAIN: return launch_col_fwd_mode<COL_PLAIN>(a, rows, s);
  }
  return cudaErrorInvalidValue;
}
}  // namespace hy
