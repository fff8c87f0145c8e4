#!/bin/bash
# The driver's scaling run, as the builder can reproduce it: for N in "$@" (default 1 2 4 8) the reference arm and the
# B200 arm of `bench.py --gpus N --steps 20 --warmup 5`, N > 1 under torch.distributed.run.  Lines land in
# gpurun_out/scale_{b200,ref}_N.json.   Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/scale_check.sh'
set -u
mkdir -p gpurun_out
NS=${@:-1 2 4 8}
port=29600
for n in $NS; do
  for impl in reference b200; do
    port=$((port + 1))
    T0=$SECONDS
    if [ "$n" = 1 ]; then
      timeout 400 python bench.py --impl $impl --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_${impl}_$n.json 2> gpurun_out/scale_${impl}_$n.err < /dev/null
    else
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
          bench.py --impl $impl --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_${impl}_$n.json 2> gpurun_out/scale_${impl}_$n.err < /dev/null
    fi
    echo "N=$n $impl rc=$? $((SECONDS - T0)) s  $(tail -n 1 gpurun_out/scale_${impl}_$n.err | cut -c1-160)"
  done
done
timeout 60 python tools/bench_summary.py $(for n in $NS; do echo gpurun_out/scale_b200_$n.json gpurun_out/scale_reference_$n.json; done) < /dev/null
