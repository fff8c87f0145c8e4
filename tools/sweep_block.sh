#!/bin/bash
# Two-stream (the headline configuration) A/B of the CTA size, repeated to see box noise.
for rep in 1 2 3; do
 for block in 256 128; do
  for env in cartpole mountain_car pendulum; do
    python bench.py --env $env --block $block --no-cpu-baseline --no-e2e --rollout-steps 0 --steps 1000 \
      | python tools/show_bench.py "$env block=$block"
  done
 done
done
