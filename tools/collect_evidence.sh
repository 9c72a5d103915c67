#!/bin/bash
# One gpurun call that produces everything profiles/ cites for a round (run from the repo root on the GPU box).
#   bash tools/collect_evidence.sh r1
set -u
R=${1:-rX}
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3 | tee $O/${R}_pytest_gpu.txt
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/${R}_smoke.txt
echo "== bench"; python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; tail -c 400 $O/${R}_bench_n1.json; echo
echo "== bench reference arm"; python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_n1_reference_arm.json 2>/dev/null; cut -c1-200 $O/${R}_bench_n1_reference_arm.json
echo "== bench 4096-point rows"; HYENA_B200_LOGM2=12 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/${R}_bench_logm2_12.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/${R}_bench_logm2_12.json')); print(d['ms_per_step'], d['roofline']['span_ms_per_step'])"
echo "== torch.fft GPU comparator"; python tests/perf_torch_fft_gpu.py --steps 3 > $O/${R}_torch_fft_gpu_comparator.json 2>/dev/null; cat $O/${R}_torch_fft_gpu_comparator.json
echo "== ncu launch list (time)"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_time.csv python tools/prof_step.py --warmup 1 --steps 1 > /dev/null 2>&1
echo "== ncu launch list (dram bytes, warm caches)"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --csv --log-file $O/${R}_launches_dram.csv python tools/prof_step.py --warmup 1 --steps 1 > /dev/null 2>&1
for k in row_pass_kernelILi3ELi10E row_pass_kernelILi1ELi10E col_inv_kernelILi10ELi10ELi1E col_fwd_kernelILi10ELi10ELi1E filter_tc_bwd_kernel filter_tc_red_kernel filter_tc_fwd_kernel; do
  echo "== ncu full $k"
  ncu --set full --cache-control none --clock-control none --import-source on --kernel-name-base mangled -k regex:$k -s 1 -c 1 -o /tmp/f_$k -f python tools/prof_step.py --warmup 1 --steps 1 > /dev/null 2>&1
  ncu -i /tmp/f_$k.ncu-rep --page raw --csv > $O/${R}_ncu_$k.raw.csv 2>/dev/null
done
cp /tmp/f_row_pass_kernelILi3ELi10E.ncu-rep $O/${R}_row_pass_bwd1.ncu-rep 2>/dev/null
ls -la $O | tail -30
