#!/bin/bash
# ncu evidence for the high-occupancy build (gymrs_set_launch_occupancy): one --set full capture of the step
# kernel per env in steady state with `bench.py --wide`, next to the default build's capture of profile_round.sh.
# Usage: gpurun --timeout 900 -- 'bash tools/profile_wide.sh'   then   python profiles/summarize.py r02 wide
set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --steps 200 --warmup 20 --wide --no-e2e --rollout-steps 0"
for env in cartpole mountain_car pendulum; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6000 -c 2 -f -o gpurun_out/prof_wide_$env \
      python bench.py --env $env $B > gpurun_out/ncu_full_wide_$env.log 2>&1 < /dev/null
  echo "full wide $env rc=$?"
done
ls -la gpurun_out/prof_wide_* | tail
