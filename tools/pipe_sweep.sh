rm -f gpurun_out/r2_ab.txt
bash tools/ab_r2.sh base
for cfg in 2,4 3,2 3,4 4,2 4,1 2,8 4,4 6,1; do bash tools/ab_r2.sh pipe_$cfg HYENA_B200_PIPE=$cfg; done
cat gpurun_out/r2_ab.txt
HYENA_B200_PIPE=3,2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "golden or fftconv or host_step" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity_full.py -x -q --timeout 600 2>&1 | tail -12
HYENA_B200_PIPE=3,2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/pipe32_dram.csv python tools/prof_step.py --warmup 1 --steps 1 > /dev/null 2>&1
ls -la gpurun_out/pipe32_dram.csv
