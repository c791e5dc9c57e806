#!/bin/bash
out=gpurun_out/r02_pdl_late.txt
for v in 0 1 2 3; do for b in 512 64; do
  echo "== PDL_LATE=$v (bit 0: conv fwd/dgrad, bit 1: wgrad) B=$b" >> $out
  HIPPIE_B200_PDL_LATE=$v B=$b STEPS=50 timeout 120 python tools/quick_bench.py 2>&1 | grep -E "train" >> $out
done; done
cat $out
