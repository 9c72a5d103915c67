#define HY_MODE 3
#include "k_col_inv.inc"
