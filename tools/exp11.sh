#!/bin/bash
# compute-sanitizer evidence: memcheck / racecheck / synccheck / initcheck on a B=32 train step + embed
export HIPPIE_B200_GRAPHS=0
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
