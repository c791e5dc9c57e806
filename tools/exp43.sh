#!/bin/bash
# NCCL streams at high priority vs torch's default (N GPUs)
N=${1:-2}
out=gpurun_out/r02_exp43_n$N.txt
{
for rep in 1 2; do for hp in 0 1; do
  HIPPIE_B200_NCCL_HIPRIO=$hp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$hp bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline > /tmp/b.json 2> /tmp/b.err || { echo "bench failed"; tail -5 /tmp/b.err; }
  python - <<PY
import json
d = json.load(open("/tmp/b.json"))
print("NCCL_HIPRIO=$hp N=$N rep $rep: %.0f samples/s  %.3f ms/step  e2e %.0f  bs64 %.3f ms  embed %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["other_workloads"]["supervised_bs64"]["ms_per_step"], d["other_workloads"]["embed"]["value"]))
PY
done; done
} > $out 2>&1
cat $out
