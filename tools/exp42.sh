#!/bin/bash
# needs a library built with: make -C hippie_b200/csrc clean all EXTRA=-DHP_EXPERIMENTS
# knock-outs (wrong results): 1 = no weight gradients, 2 = no BatchNorm-backward reduce launches, 4 = no BatchNorm-backward apply launches
out=gpurun_out/r02_exp42.txt
{
for rep in 1 2; do for sk in 0 2 1 3; do
  echo "== DEBUG_SKIP=$sk rep $rep"
  HIPPIE_B200_DEBUG_SKIP=$sk B=512 STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train"
done; done
} > $out 2>&1
cat $out
