#!/bin/bash
# which GEMM kinds gain from the fused N = 2 BN instruction on the step; weight planes kept current by AdamW vs converted per call
out=gpurun_out/r02_exp32.txt
{
for rep in 1 2; do
for sch in 0 7 1 3 5; do
  echo "== MMA_SCHEME=$sch rep $rep"; HIPPIE_B200_MMA_SCHEME=$sch B=512 STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train"
done
done
for keep in 0 1; do for b in 512 64; do
  echo "== KEEP_PLANES=$keep B=$b"; HIPPIE_B200_KEEP_PLANES=$keep B=$b STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"
done; done
} > $out 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $out
cat $out
