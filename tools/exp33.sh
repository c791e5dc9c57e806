#!/bin/bash
# tools/libold.so / libB.so / lib4.so / libprev.so: libraries built from the commit before (or with the variant named in the echo lines) and copied next to this script
# register budgets: GEMM CTAs at 96 registers (A: BatchNorm kernels capped at 112 too, B: BatchNorm kernels uncapped) vs the
# round-2 library (old: 122 registers, three N = 64 MMAs per k-step)
out=gpurun_out/r02_exp33.txt
cp hippie_b200/libhippie_b200.so /tmp/A.so
{
for rep in 1 2; do
  for which in old A B; do
    case $which in old) cp tools/libold.so hippie_b200/libhippie_b200.so;; A) cp /tmp/A.so hippie_b200/libhippie_b200.so;; B) cp tools/libB.so hippie_b200/libhippie_b200.so;; esac
    for b in 512 64; do echo "== $which B=$b rep $rep"; B=$b STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"; done
  done
done
cp /tmp/A.so hippie_b200/libhippie_b200.so
} > $out 2>&1
python -m pytest tests/test_gpu_parity2.py -m gpu -x -q -k "planes or behind" 2>&1 | tail -5 >> $out
cat $out
