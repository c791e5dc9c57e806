#!/bin/bash
out=gpurun_out/r02_wgrad_minkb.txt
for kb in 4 8 16 32 64; do for b in 512 64; do
  echo "== WGRAD_MIN_KB=$kb B=$b" >> $out
  HIPPIE_B200_WGRAD_MIN_KB=$kb B=$b STEPS=50 timeout 120 python tools/quick_bench.py 2>&1 | grep -E "train" >> $out
done; done
cat $out
