#!/bin/bash
# round-2 experiment 14: NCCL CTA budget for the overlapped gradient exchange (N GPUs of one box)
N=${1:-2}
out=gpurun_out/r02_nccl_ctas_n$N.txt
for ctas in default 2 4 8 16; do
  if [ $ctas = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline > /tmp/b.json 2> /tmp/b.err
  python - <<PY >> $out
import json
d = json.load(open("/tmp/b.json"))
print("NCCL_MAX_CTAS=$ctas N=$N: %.0f samples/s  %.3f ms/step  e2e %.0f  bs64 %.3f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["other_workloads"]["supervised_bs64"]["ms_per_step"]))
PY
done
cat $out
