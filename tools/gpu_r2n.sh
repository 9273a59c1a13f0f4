#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "sweep or alignment or fuzz or mixed or fused_clock_every or fused_equals" > $out/tests.log 2>&1; echo "rc=$?" >> $out/tests.log; tail -3 $out/tests.log
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 2>> $out/err.log | python tools/benchline.py "$wl $tag"; }
for wl in w1500 w750 w375; do
  run "no rotation" $wl AFSK_NO_ROT=1
  run "rotated" $wl X=1
  run "no rotation" $wl AFSK_NO_ROT=1
  run "rotated" $wl X=1
done
