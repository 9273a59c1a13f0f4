#!/bin/bash
out=gpurun_out/r2m; mkdir -p $out
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 2>> $out/err.log | python tools/benchline.py "$wl $tag"; }
for pair in "c2 40" "w600 80" "c4 160" "w1000 48" "w500 96"; do
  set -- $pair
  run base $1 X=1
  run "shift<$2,1>" $1 AFSK_SHIFT_BFS=$2
  run base $1 X=1
  run "shift<$2,1>" $1 AFSK_SHIFT_BFS=$2
done
tail -3 $out/err.log
