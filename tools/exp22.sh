#!/bin/bash
N=${1:-8}
out=gpurun_out/r02_scale_n$N.txt
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > /tmp/b1.json 2> /tmp/b1.err
python - <<PY >> $out
import json
d = json.load(open("/tmp/b1.json"))
print("N=1: %.0f samples/s  %.3f ms/step  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
for mode in 0 1; do
  HIPPIE_B200_DP_PARTS=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29630 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline > /tmp/b.json 2> /tmp/b.err || { echo "bench failed (parts=$mode)" >> $out; tail -5 /tmp/b.err >> $out; }
  [ $mode = 0 ] && cp /tmp/b.json gpurun_out/r02_bench_${N}gpu.json
  python - <<PY >> $out
import json
try:
    d = json.load(open("/tmp/b.json"))
    print("DP_PARTS=$mode N=$N: %.0f samples/s  %.3f ms/step  e2e %.0f  bs64 %.3f ms  embed %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["other_workloads"]["supervised_bs64"]["ms_per_step"], d["other_workloads"]["embed"]["value"]))
except Exception as e:
    print("no result", e)
PY
done
cat $out
