#!/bin/bash
# Persistent TMA-staged step kernel (vec = 8) on ONE stream, chained: ring depth x CTAs per SM.
for stages in 2 3 4; do
  GYMRS_NVCC_EXTRA="-DGYMRS_STREAM_STAGES=$stages" python -c "from gym_rs_b200 import build; build.build(force=True)"
  for per_sm in 1 2 3 4; do
    GYMRS_STREAM_CTAS_PER_SM=$per_sm python bench.py --streams 1 --pdl 2 --vec 8 --no-cpu-baseline --no-e2e --rollout-steps 0 --steps 1000 \
      | python tools/show_bench.py "stages=$stages ctas/sm=$per_sm"
  done
done
python -c "from gym_rs_b200 import build; build.build(force=True)"
