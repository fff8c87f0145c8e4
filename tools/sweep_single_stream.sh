#!/bin/bash
# One-stream launch-geometry sweep on the cold ring (main figure = `value` with --streams 1).
for pdl in 2 1; do
 for block in 64 128 256; do
  for vec in 4 2; do
    python bench.py --streams 1 --pdl $pdl --block $block --vec $vec --no-cpu-baseline --no-e2e --rollout-steps 0 --steps 1000 \
      | python tools/show_bench.py "pdl=$pdl block=$block vec=$vec"
  done
 done
done
