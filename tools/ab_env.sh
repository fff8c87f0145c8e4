#!/bin/bash
# Interleaved A/B of ENVIRONMENT settings on one box:  tools/ab_env.sh <env> <rounds> "<VAR=val ...>" "<VAR=val ...>" ...
# (use "-" for no setting).  Same summary line as tools/ab_interleaved.sh.
env=$1; rounds=$2; shift 2
for r in $(seq 1 $rounds); do
  i=0
  for spec in "$@"; do
    i=$((i+1))
    if [ "$spec" = "-" ]; then pre=""; else pre="$spec"; fi
    env $pre python bench.py --env $env --steps 2000 --warmup 50 --no-cpu-baseline --no-e2e --rollout-steps 0 > gpurun_out/abe_${env}_${i}_$r.json 2> gpurun_out/abe_${env}_${i}_$r.err || tail -3 gpurun_out/abe_${env}_${i}_$r.err
    python - "$spec" gpurun_out/abe_${env}_${i}_$r.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
us = lambda x: x["ms_per_step"] * 1e3
s = d["single_stream_default"]
print("%-26s 2-stream %.3f  pdl1 cold %.3f res %.3f  pdl2 cold %.3f res %.3f  clk %s" % (
    sys.argv[1], us(d), us(s["cold_ring"]), us(s["l2_resident"]), us(d["single_stream_chained"]), us(d["l2_resident"]), d["clocks"]["sm_mhz"]))
PY
  done
done
