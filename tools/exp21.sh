#!/bin/bash
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15
for mode in 1 0; do
  HIPPIE_B200_DP_PARTS=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline > /tmp/b.json 2> /tmp/b.err || { echo "bench failed (parts=$mode)"; tail -5 /tmp/b.err; }
  python - <<PY
import json
try:
    d = json.load(open("/tmp/b.json"))
    print("DP_PARTS=$mode N=$N: %.0f samples/s  %.3f ms/step  e2e %.0f  bs64 %.3f ms  loss %.5f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["other_workloads"]["supervised_bs64"]["ms_per_step"], d["loss_last"]))
except Exception as e:
    print("no result", e)
PY
done
