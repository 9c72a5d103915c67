#define HY_MODE 2
#include "k_col_inv.inc"
