#!/bin/bash
# round-2 experiment 8: per-launch choice of the GEMM variant (2-stage / two CTAs per SM vs 3-stage / one CTA per SM)
out=gpurun_out/r02_variants_step.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $out
for pv in 1 0 2; do for wv in 1 2; do for b in 512 64; do
  echo "== PAIR_VARIANT=$pv WGRAD_VARIANT=$wv B=$b" >> $out
  HIPPIE_B200_PAIR_VARIANT=$pv HIPPIE_B200_WGRAD_VARIANT=$wv B=$b STEPS=50 timeout 120 python tools/quick_bench.py 2>&1 | grep -E "train|embed" >> $out
done; done; done
cat $out
