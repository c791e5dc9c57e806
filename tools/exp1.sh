#!/bin/bash
# round-2 experiment 1: ingest ceiling + tile-shape variants of the pair GEMMs
cd tools
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 120 ./tma_bw > ../gpurun_out/r02_tma_bw.txt 2>&1
for v in pair_test pair_test_64x3 pair_test_128x2 pair_test_128x3; do
  echo "=== $v" >> ../gpurun_out/r02_pair_variants.txt
  PT_STAMPS=1 timeout 300 ./$v 512 1 2>&1 | grep -v "^$" >> ../gpurun_out/r02_pair_variants.txt
  timeout 300 ./$v 512 2 2>&1 >> ../gpurun_out/r02_pair_variants.txt
  timeout 300 ./$v 512 3 2>&1 >> ../gpurun_out/r02_pair_variants.txt
  timeout 300 ./$v 512 4 2>&1 >> ../gpurun_out/r02_pair_variants.txt
done
tail -5 ../gpurun_out/r02_tma_bw.txt
