#!/bin/bash
# ncu --set full of the round-2 kernels at the bench shape (one capture per kernel), raw + source pages exported as CSV
set -u
O=gpurun_out; mkdir -p $O
KERNELS=${@:-wgrad_kernel proj_gemm_kernel row_pass_bwd1_staged_kernel row_pass_fwd_staged_kernel col_inv_kernelILi10ELi10ELi1E col_inv_kernelILi10ELi10ELi0E col_fwd_kernelILi10ELi10ELi1E filter_tc_bwd_kernel filter_tc_red_kernel filter_tc_fwd2_kernel}
for k in $KERNELS; do
  timeout 300 ncu --set full --cache-control none --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 1 -c 1 -o /tmp/f_$k -f python tools/prof_step.py --warmup 1 --steps 1 > /dev/null 2>&1
  ncu -i /tmp/f_$k.ncu-rep --page raw --csv > $O/r2_ncu_$k.raw.csv 2>/dev/null
  ncu -i /tmp/f_$k.ncu-rep --page source --csv > $O/r2_ncu_$k.source.csv 2>/dev/null
  echo "$k done: $(wc -c < $O/r2_ncu_$k.raw.csv) bytes"
done
