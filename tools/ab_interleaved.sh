#!/bin/bash
# Interleaved A/B of library variants on ONE box (drift and power capping hit all variants alike):
#   tools/ab_interleaved.sh <env> <rounds> <variant>[@<bench args, + for spaces>] ...   ("default" = the in-tree library)
# e.g.  tools/ab_interleaved.sh cartpole 3 old merge merge@--block+256
env=$1; rounds=$2; shift 2
for r in $(seq 1 $rounds); do
  for spec in "$@"; do
    v=${spec%%@*}; extra=""; tag=$v
    if [ "$spec" != "$v" ]; then extra=$(echo "${spec#*@}" | tr '+' ' '); tag=$(echo "$spec" | tr -c 'A-Za-z0-9_\n' '_'); fi
    if [ "$v" = default ]; then unset GYMRS_LIB_PATH; else export GYMRS_LIB_PATH=$PWD/gym_rs_b200/variants/libgymrs_b200_$v.so; fi
    python bench.py --env $env --steps 2000 --warmup 50 --no-cpu-baseline --no-e2e --rollout-steps 0 $extra > gpurun_out/abi_${env}_${tag}_$r.json 2> gpurun_out/abi_${env}_${tag}_$r.err || tail -3 gpurun_out/abi_${env}_${tag}_$r.err
    python - "$spec" gpurun_out/abi_${env}_${tag}_$r.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
us = lambda x: x["ms_per_step"] * 1e3
s = d["single_stream_default"]
print("%-22s 2-stream %.3f  pdl1 cold %.3f res %.3f  pdl2 cold %.3f res %.3f  clk %s" % (
    sys.argv[1], us(d), us(s["cold_ring"]), us(s["l2_resident"]), us(d["single_stream_chained"]), us(d["l2_resident"]), d["clocks"]["sm_mhz"]))
PY
  done
done
