#!/bin/bash
# One GPU-box pass: parity tests through the C ABI, smoke(), and a default bench line per env.
# Usage (from the repo root): gpurun --timeout 900 -- 'bash tools/gpu_check.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 600 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
for env in ${GYMRS_CHECK_ENVS:-cartpole}; do
  timeout 300 python bench.py --env $env > gpurun_out/bench_$env.json 2> gpurun_out/bench_$env.err
  echo "bench $env rc=$?"
  python tools/show_bench.py $env < gpurun_out/bench_$env.json
done
