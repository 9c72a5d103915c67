#!/bin/bash
# A/B runs of kernel variants on one B200 (ms per step and per-kernel ms): usage tools/ab_r2.sh <tag> [VAR=VALUE ...]
# each invocation appends one line to gpurun_out/r2_ab.txt
tag=$1; shift
out=gpurun_out/r2_ab.txt
mkdir -p gpurun_out
line=$(env "$@" python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference 2> gpurun_out/ab_${tag}.err | tail -1)
python - "$tag" "$*" <<PY >> $out
import json, sys
tag, envs = sys.argv[1], sys.argv[2]
try:
    b = json.loads('''$line''')
    r = b["roofline"]
    ks = " ".join(f"{k}={v['ms_per_step']:.2f}" for k, v in r["kernels"].items() if v["ms_per_step"] > 0.05)
    print(f"{tag:28s} step {b['ms_per_step']:.2f} ms  span {r['span_ms_per_step']:.2f}  sm_mhz {b['clocks']['sm_mhz']}  [{envs}]  {ks}")
except Exception as e:
    print(f"{tag:28s} FAILED {e!r} [{envs}]")
PY
