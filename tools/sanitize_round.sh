#!/bin/bash
# compute-sanitizer passes behind profiles/r0N_sanitizer_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small ok' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
