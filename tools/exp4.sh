#!/bin/bash
cd tools
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for v in pair_test pair_test_64x3 pair_test_64x4 pair_test_128x3; do
  echo "=== $v" >> ../gpurun_out/r02_multiprod_stages.txt
  for sel in 1 3; do PT_STAMPS=1 timeout 300 ./$v 512 $sel 2>&1 | grep -v "^$" >> ../gpurun_out/r02_multiprod_stages.txt; done
done
