#!/bin/bash
# priority assignments (chain 0 = wave branch, chain 1 = ISI branch, weight gradients); 0 = least, -5 = greatest
out=gpurun_out/r02_exp41.txt
{
for rep in 1 2; do for pr in "0,0,-5" "0,-2,-5" "-2,0,-5" "0,0,-1" "-1,-1,0" "0,-5,-5" "0,0,0"; do
  echo "== PRIO=$pr rep $rep"
  HIPPIE_B200_PRIO=$pr B=512 STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train"
done; done
} > $out 2>&1
cat $out
