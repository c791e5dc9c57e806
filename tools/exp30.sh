#!/bin/bash
# tools/libold.so / libB.so / lib4.so / libprev.so: libraries built from the commit before (or with the variant named in the echo lines) and copied next to this script
# needs a library built with: make -C hippie_b200/csrc clean all EXTRA=-DHP_EXPERIMENTS
# A/B of the concatenated-B MMA scheme: old (three N = 64 MMAs per k-step) vs new (N = 128 + N = 64)
out=gpurun_out/r02_concat_ab.txt
{
echo "== pair_test OLD"; ./tools/pair_test_old 512 2>&1 | tail -40
echo "== pair_test NEW"; ./tools/pair_test 512 2>&1 | tail -40
for b in 512 64; do echo "== quick_bench NEW B=$b"; B=$b STEPS=200 python tools/quick_bench.py 2>&1 | tail -3; done
echo "== BWD_MMA=2 (speed only)"; HIPPIE_B200_BWD_MMA=2 B=512 STEPS=200 python tools/quick_bench.py 2>&1 | tail -3
} > $out 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -x -q 2>&1 | tail -5 >> $out
cat $out
