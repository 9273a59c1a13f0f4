#!/bin/bash
out=gpurun_out/r2e3; mkdir -p $out
run() { # tag workload envs...
  tag=$1; wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/b_${wl}_$tag.json 2>> $out/err.log
  python tools/benchline.py "$wl $tag" < $out/b_${wl}_$tag.json | tee -a $out/summary.txt
}
for wl in c3 w3000 w2000; do run base $wl X=1; run tile48 $wl AFSK_LANE_TILE48=1; run base $wl X=1; run tile48 $wl AFSK_LANE_TILE48=1; done
for wl in w4000 w2400; do run nsub1 $wl X=1; run nsub2 $wl AFSK_DEMOD_NSUB=2; run nsub1 $wl X=1; run nsub2 $wl AFSK_DEMOD_NSUB=2; done
AFSK_LANE_TILE48=1 AFSK_DEMOD_NSUB=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or alignment or fuzz or mixed" > $out/tests.log 2>&1; echo "rc=$?" >> $out/tests.log; tail -3 $out/tests.log
