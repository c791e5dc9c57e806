#!/bin/bash
# round-2 (second half) evidence: bench line, ncu launch list (warm caches), ncu --set full of conv_pair_kernel, CUPTI timeline
set -x
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err || exit 1
HIPPIE_B200_GRAPHS=0 python tools/profile_step.py > gpurun_out/r02b_profile_step.log 2>&1 || exit 1
HIPPIE_B200_GRAPHS=0 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02b_launches_warm.csv python tools/profile_step.py > gpurun_out/r02b_ncu1.log 2>&1
HIPPIE_B200_GRAPHS=0 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 14 -c 3 \
  -o gpurun_out/r02b_conv_pair python tools/profile_step.py > gpurun_out/r02b_ncu2.log 2>&1
ncu -i gpurun_out/r02b_conv_pair.ncu-rep --page raw --csv > gpurun_out/r02b_conv_pair_ncu_raw.csv 2>/dev/null
python tools/trace_step.py gpurun_out/r02b_trace.json > gpurun_out/r02b_timeline.txt 2>&1
rm -f gpurun_out/r02b_trace.json
tail -3 gpurun_out/r02b_ncu2.log; head -c 600 gpurun_out/r02b_bench.json
