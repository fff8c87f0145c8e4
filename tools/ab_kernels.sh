#!/bin/bash
# A/B of library variants on ONE box: tools/ab_kernels.sh <env> <variant.so> [<variant.so> ...]
# ("default" = the in-tree library).  Prints one summary block per variant.
env=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset GYMRS_LIB_PATH; else export GYMRS_LIB_PATH=$PWD/gym_rs_b200/variants/libgymrs_b200_$v.so; fi
  for rep in 1 2; do
    python bench.py --env $env --steps 500 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/ab_${env}_${v}_$rep.json 2> gpurun_out/ab_${env}_${v}_$rep.err || tail -3 gpurun_out/ab_${env}_${v}_$rep.err
    python tools/bench_summary.py gpurun_out/ab_${env}_${v}_$rep.json
  done
done
