#!/bin/bash
# tools/libold.so / libB.so / lib4.so / libprev.so: libraries built from the commit before (or with the variant named in the echo lines) and copied next to this script
# GEMM CTAs at 96 (cur) vs 80 (lib4) registers
out=gpurun_out/r02_exp34.txt
cp hippie_b200/libhippie_b200.so /tmp/cur.so
{
for rep in 1 2; do
  for which in cur lib4; do
    case $which in cur) cp /tmp/cur.so hippie_b200/libhippie_b200.so;; lib4) cp tools/lib4.so hippie_b200/libhippie_b200.so;; esac
    for b in 512 64; do echo "== $which B=$b rep $rep"; B=$b STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"; done
  done
done
cp /tmp/cur.so hippie_b200/libhippie_b200.so
} > $out 2>&1
python -m pytest tests/test_gpu_parity2.py -m gpu -x -q -k "planes or behind" 2>&1 | tail -5 >> $out
cat $out
