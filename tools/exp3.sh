#!/bin/bash
cd tools
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for v in pair_test_old pair_test; do
  echo "=== $v" >> ../gpurun_out/r02_multiprod.txt
  for sel in 1 2 3 4; do PT_STAMPS=1 timeout 300 ./$v 512 $sel 2>&1 | grep -v "^$" >> ../gpurun_out/r02_multiprod.txt; done
done
grep -c "pair" ../gpurun_out/r02_multiprod.txt
