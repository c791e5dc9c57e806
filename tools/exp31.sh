#!/bin/bash
# tools/libold.so / libB.so / lib4.so / libprev.so: libraries built from the commit before (or with the variant named in the echo lines) and copied next to this script
# (1) old vs new library on the bs512 / bs64 step, interleaved; (2) weight-gradient CTAs limited to one per SM by a
# shared-memory pad; (3) gradient error table with the (gradient lo) x (hi) product dropped in dgrad / wgrad
out=gpurun_out/r02_exp31.txt
cp hippie_b200/libhippie_b200.so /tmp/new.so
{
for rep in 1 2; do
  for which in old new; do
    [ $which = old ] && cp tools/libold.so hippie_b200/libhippie_b200.so || cp /tmp/new.so hippie_b200/libhippie_b200.so
    for b in 512 64; do echo "== $which B=$b rep $rep"; B=$b STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train|^embed"; done
  done
done
cp /tmp/new.so hippie_b200/libhippie_b200.so
for pad in 20 0; do for kb in 16 8; do
  echo "== WGRAD_PAD_KB=$pad MIN_KB=$kb"; HIPPIE_B200_WGRAD_PAD_KB=$pad HIPPIE_B200_WGRAD_MIN_KB=$kb B=512 STEPS=300 python tools/quick_bench.py 2>&1 | grep -E "^train"
done; done
echo "== grad table BWD_MMA=2"
HIPPIE_B200_BWD_MMA=2 python tools/grad_table.py gpurun_out/r02_grad_error_table_bwd2mma.md 2>&1 | tail -2
} > $out 2>&1
cat $out
