#!/bin/bash
# round-2 evidence run: GPU tests, bench (N=1, both arms), launch list, full capture of the K=1536 conv launch, timeline
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_full.log 2>&1; tail -4 gpurun_out/r02_gputest_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
HIPPIE_B200_GRAPHS=0 python tools/profile_step.py > gpurun_out/r02_profile_step_plain.log 2>&1 && \
HIPPIE_B200_GRAPHS=0 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_warm.csv python tools/profile_step.py > gpurun_out/r02_ncu_warm.log 2>&1
HIPPIE_B200_GRAPHS=0 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 14 -c 3 \
    -f -o gpurun_out/r02_conv_pair python tools/profile_step.py > gpurun_out/r02_ncu_full.log 2>&1
python tools/trace_step.py gpurun_out/r02_trace_step.json > gpurun_out/r02_timeline.txt 2>&1
rm -f gpurun_out/r02_trace_step.json
tail -c 400 gpurun_out/r02_bench.json; tail -c 600 gpurun_out/r02_bench_reference.json
