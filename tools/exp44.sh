#!/bin/bash
# 8-GPU bench line of the current build (one run, default settings)
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02b_bench_${N}gpu.json 2> gpurun_out/r02b_bench_${N}gpu.err || { echo "bench failed"; tail -5 gpurun_out/r02b_bench_${N}gpu.err; }
python - <<PY
import json
d = json.load(open("gpurun_out/r02b_bench_${N}gpu.json"))
print("N=$N: %.0f samples/s  %.3f ms/step  e2e %.0f  bs64 %.3f ms  embed %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["other_workloads"]["supervised_bs64"]["ms_per_step"], d["other_workloads"]["embed"]["value"]))
PY
