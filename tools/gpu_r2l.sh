#!/bin/bash
out=gpurun_out/r2l; mkdir -p $out
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 2>> $out/err.log | python tools/benchline.py "$wl $tag"; }
for wl in c4 c2 w600; do
  run base $wl X=1
  run "one_cta nsub4 stages2" $wl AFSK_DEMOD_ONE_CTA=2 AFSK_DEMOD_NSUB=4
  run "one_cta nsub3 stages3" $wl AFSK_DEMOD_ONE_CTA=3 AFSK_DEMOD_NSUB=3
  run base $wl X=1
  run "one_cta nsub4 stages2" $wl AFSK_DEMOD_ONE_CTA=2 AFSK_DEMOD_NSUB=4
  run "one_cta nsub2 stages4" $wl AFSK_DEMOD_ONE_CTA=4 AFSK_DEMOD_NSUB=2
done
