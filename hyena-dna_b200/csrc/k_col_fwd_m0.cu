#define HY_MODE 0
#include "k_col_fwd.inc"
