#!/bin/bash
# needs a library built with: make -C hippie_b200/csrc clean all EXTRA=-DHP_EXPERIMENTS
# floor of the step without GEMM main loops: every conv / dgrad / wgrad CTA processes at most FAKE_K k-blocks (wrong results)
out=gpurun_out/r02_exp36.txt
{
for fk in 0 4 2; do for b in 512 64; do
  echo "== FAKE_K=$fk B=$b"
  HIPPIE_B200_FAKE_K=$fk B=$b STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"
done; done
echo "== FAKE_K=2 no wgrad"
HIPPIE_B200_DEBUG_SKIP=1 HIPPIE_B200_FAKE_K=2 B=512 STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"
echo "== phases FAKE_K=2"; HIPPIE_B200_FAKE_K=2 python tools/phase_bench.py 2>&1 | tail -8
echo "== phases"; python tools/phase_bench.py 2>&1 | tail -8
} > $out 2>&1
cat $out
