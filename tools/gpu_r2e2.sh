#!/bin/bash
out=gpurun_out/r2e2; mkdir -p $out
run() { # workload nsub stages
  if [ "$2" = "-" ]; then unset AFSK_DEMOD_NSUB; else export AFSK_DEMOD_NSUB=$2; fi
  if [ "$3" = "-" ]; then unset AFSK_DEMOD_STAGES; else export AFSK_DEMOD_STAGES=$3; fi
  timeout 300 python bench.py --workload $1 --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/b_$1_$2_$3.json 2>> $out/err.log
  python tools/benchline.py "$1 nsub=$2 stages=$3" < $out/b_$1_$2_$3.json | tee -a $out/summary.txt
}
for wl in c2 c4; do run $wl 1 -; run $wl 2 -; run $wl 1 -; run $wl 2 -; run $wl 3 -; run $wl 2 3; done
for wl in w1500 w750 w375; do run $wl 2 -; run $wl 3 -; run $wl 4 -; done
for wl in w1000 w500 w800 w480 w400 w240; do run $wl 1 -; run $wl 2 -; done
