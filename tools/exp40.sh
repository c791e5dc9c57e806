#!/bin/bash
# per-node priorities in the replayed graph: weight gradients ABOVE the chains
out=gpurun_out/r02_exp40.txt
{
for rep in 1 2; do for mode in "0 0" "1 1" "1 0"; do set -- $mode; for b in 512 64; do
  echo "== GRAPH_PRIO=$1 PRIO_INVERT=$2 B=$b rep $rep"
  HIPPIE_B200_GRAPH_PRIO=$1 HIPPIE_B200_PRIO_INVERT=$2 B=$b STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train"
done; done; done
for inv in 0 1; do echo "== eager (GRAPHS=0) PRIO_INVERT=$inv"; HIPPIE_B200_GRAPHS=0 HIPPIE_B200_PRIO_INVERT=$inv B=512 STEPS=100 python tools/quick_bench.py 2>&1 | grep -E "^train"; done
} > $out 2>&1
cat $out
