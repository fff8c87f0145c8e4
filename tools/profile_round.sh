#!/bin/bash
# The ncu passes behind profiles/ (see profiles/r01_SUMMARY.md): launch list of the bench command,
# one --set full capture per step kernel in steady state, and single-pass steady-state DRAM bytes.
# Usage: gpurun --timeout 1500 -- 'bash tools/profile_round.sh'   then   python profiles/summarize.py r01
set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --steps 200 --warmup 20"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 5500 -c 600 --csv \
    --log-file gpurun_out/launches_cartpole.csv python bench.py $B --e2e-steps 2 --rollout-steps 8 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for env in cartpole mountain_car pendulum; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6000 -c 2 -f -o gpurun_out/prof_$env \
      python bench.py --env $env $B --no-e2e --rollout-steps 0 > gpurun_out/ncu_full_$env.log 2>&1
  echo "full $env rc=$?"
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none \
      --clock-control none -k regex:step_kernel -s 6000 -c 128 --csv --log-file gpurun_out/steady_dram_$env.csv \
      python bench.py --env $env $B --no-e2e --rollout-steps 0 > gpurun_out/ncu_steady_$env.log 2>&1
  echo "steady $env rc=$?"
done
ls -la gpurun_out | tail -20
