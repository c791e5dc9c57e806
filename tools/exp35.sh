#!/bin/bash
# needs a library built with: make -C hippie_b200/csrc clean all EXTRA=-DHP_EXPERIMENTS
# upper bound of what re-using one A slab for the three taps of a k3 conv / dgrad would give: A fetched for one k-block in
# three (results are wrong, speed only)
out=gpurun_out/r02_exp35.txt
{
for rep in 1 2; do
for fr in 0 1; do
  echo "== FAKE_REUSE=$fr rep $rep"
  if [ $fr = 1 ]; then export HIPPIE_B200_FAKE_REUSE=1; else unset HIPPIE_B200_FAKE_REUSE; fi
  B=512 STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"
done; done
} > $out 2>&1
cat $out
