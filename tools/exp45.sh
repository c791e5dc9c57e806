#!/bin/bash
# tools/libold.so / libB.so / lib4.so / libprev.so: libraries built from the commit before (or with the variant named in the echo lines) and copied next to this script
# eval-mode epilogue: four rows per thread and iteration (cur) vs one (prev = tools/libprev.so, built from the commit before)
out=gpurun_out/r02_exp45.txt
cp hippie_b200/libhippie_b200.so /tmp/cur.so
{
for rep in 1 2; do for which in prev cur; do
  [ $which = prev ] && cp tools/libprev.so hippie_b200/libhippie_b200.so || cp /tmp/cur.so hippie_b200/libhippie_b200.so
  for b in 4096 512 64; do echo "== $which rep $rep"; B=$b HIPPIE_ROOT=$PWD python tools/embed_bench.py 2>&1 | tail -1; done
done; done
cp /tmp/cur.so hippie_b200/libhippie_b200.so
} > $out 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -3 >> $out
HIPPIE_B200_GRAPHS=0 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_embed_launches.csv python tools/embed_once.py > gpurun_out/embed_ncu.log 2>&1; python tools/launch_summary.py gpurun_out/r02c_embed_launches.csv --grid | head -12 >> $out
cat $out
