#!/bin/bash
# A/B the experimental kernel variants on one B200 (run through gpurun): for each flag setting, the parity subset that
# exercises the long-convolution core, then a short bench line without the e2e / CPU legs.  Results -> gpurun_out/ab_*.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/ab_flags.sh'
set -u
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 120 python -m pytest tests -m gpu -x -q --timeout 100 \
      -k "golden or fp64 or host_step or large_1m_sampled" > gpurun_out/ab_${tag}_pytest.txt 2>&1
  tail -1 gpurun_out/ab_${tag}_pytest.txt
  env "$@" timeout 60 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline \
      > gpurun_out/ab_${tag}.json 2> gpurun_out/ab_${tag}.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{tag}.json"))
    k = d["roofline"]["kernels"]
    print(tag, "step %.2f ms  span %.2f ms " % (d["ms_per_step"], d["roofline"]["span_ms_per_step"]),
          {n: round(v["ms_per_step"], 2) for n, v in k.items() if v["ms_per_step"] > 0.5}, d["clocks"]["sm_mhz"])
except Exception as e:   # a failed variant must not hide the others
    print(tag, "FAILED:", e)
PY
}
run default          HYENA_B200_NOOP=1
run bwd1_3cta        HYENA_B200_ROW_BWD1_CTAS=3
run bwd1_staged      HYENA_B200_ROW_BWD1_STAGE=1
run fwd_staged       HYENA_B200_ROW_FWD_STAGE=1
run both_staged      HYENA_B200_ROW_BWD1_STAGE=1 HYENA_B200_ROW_FWD_STAGE=1
run fused_coop       HYENA_B200_FUSED=1 HYENA_B200_FUSED_MB=2048
run fused_flow       HYENA_B200_FUSED=2
