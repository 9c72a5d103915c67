#define HY_MODE 1
#include "k_col_inv.inc"
