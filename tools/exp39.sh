#!/bin/bash
# per-node priorities in the replayed graph (chains high, weight gradients low)
out=gpurun_out/r02_exp39.txt
{
for rep in 1 2; do for pr in 0 1; do for b in 512 64; do
  echo "== GRAPH_PRIO=$pr B=$b rep $rep"
  HIPPIE_B200_GRAPH_PRIO=$pr B=$b STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"
done; done; done
} > $out 2>&1
python -m pytest tests/test_gpu_parity2.py -m gpu -x -q -k "helper or planes or behind" 2>&1 | tail -4 >> $out
cat $out
