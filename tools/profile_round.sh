#!/bin/bash
# The ncu passes behind profiles/ (see profiles/r02_SUMMARY.md): per env the launch list of the bench
# command, one --set full capture of the step kernel in steady state and the single-pass steady-state DRAM
# bytes; plus one --set full capture of the fused rollout kernel (CartPole).
# Usage: gpurun --timeout 1500 -- 'bash tools/profile_round.sh'   then   python profiles/summarize.py r02
set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --steps 200 --warmup 20"
for env in cartpole mountain_car pendulum; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 5500 -c 600 --csv \
      --log-file gpurun_out/launches_$env.csv python bench.py --env $env $B --e2e-steps 2 --rollout-steps 8 > gpurun_out/ncu_launches_$env.log 2>&1
  echo "launch list $env rc=$?"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6000 -c 2 -f -o gpurun_out/prof_$env \
      python bench.py --env $env $B --no-e2e --rollout-steps 0 > gpurun_out/ncu_full_$env.log 2>&1
  echo "full $env rc=$?"
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none \
      --clock-control none -k regex:step_kernel -s 6000 -c 128 --csv --log-file gpurun_out/steady_dram_$env.csv \
      python bench.py --env $env $B --no-e2e --rollout-steps 0 > gpurun_out/ncu_steady_$env.log 2>&1
  echo "steady $env rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -f -o gpurun_out/prof_rollout_cartpole \
    python bench.py $B --no-e2e --burn-in 20 --rollout-steps 64 > gpurun_out/ncu_full_rollout.log 2>&1
echo "full rollout rc=$?"
ls -la gpurun_out | tail -20
