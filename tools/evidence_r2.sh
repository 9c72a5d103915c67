#!/bin/bash
# Round-2 evidence in one gpurun call (run from the repo root on the GPU box): tests, smoke, bench (own + reference arm),
# launch list with time and DRAM bytes, full ncu captures of the dominant kernels.  Outputs under gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | tee $O/r2_pytest_gpu.txt
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r2_smoke.txt
echo "== bench"; python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 600 $O/r2_bench_n1.json; echo
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --csv --log-file $O/r2_launches.csv python tools/prof_step.py --warmup 1 --steps 1 > /dev/null 2>&1
python tools/launch_summary.py $O/r2_launches.csv --steps 1 --skip-steps 1 | tee $O/r2_launches_summary.txt | tail -5
bash tools/ncu_r2.sh
for f in $O/r2_ncu_*.raw.csv; do python tools/ncu_summary.py $f > ${f%.raw.csv}.txt; done
echo "== LN"; python tools/bench_ln.py | tee $O/r2_bench_ln.json
ls $O | head -60
