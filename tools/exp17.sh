#!/bin/bash
out=gpurun_out/r02_knockout.txt
for v in 0 1; do for b in 512; do
  echo "== DEBUG_SKIP=$v (1: no conv weight-gradient launches) B=$b" >> $out
  HIPPIE_B200_DEBUG_SKIP=$v B=$b STEPS=50 timeout 120 python tools/quick_bench.py 2>&1 | grep -E "train|launches" >> $out
done; done
B=512 timeout 200 python tools/phase_bench.py >> $out 2>&1
cat $out
