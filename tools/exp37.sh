#!/bin/bash
# shortcut convolutions on a helper stream; gradient memset under the forward pass
out=gpurun_out/r02_exp37.txt
{
for rep in 1 2; do for sh in 0 1; do for b in 512 64; do
  echo "== SHORTCUT_STREAM=$sh B=$b rep $rep"
  HIPPIE_B200_SHORTCUT_STREAM=$sh B=$b STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"
done; done; done
} > $out 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -5 >> $out
cat $out
