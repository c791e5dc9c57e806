#!/bin/bash
# round-2 experiment 7: MMA issuer as a warp-uniform loop with an elected lane (straight-line UTCHMMA)
cd tools
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
out=../gpurun_out/r02_elect.txt
for v in pair_test_r02a pair_test pair_test_64x3 pair_test_64x4 pair_test_128x2 pair_test_128x3; do
  echo "=== $v" >> $out
  for sel in 1 2 3 4; do PT_STAMPS=1 timeout 300 ./$v 512 $sel 2>&1 | grep -v "^$" >> $out; done
done
grep -c pair $out
