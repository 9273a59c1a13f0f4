#!/bin/bash
out=gpurun_out/r2b5; mkdir -p $out
AFSK_FUSED=2 timeout 1500 python -m pytest tests -m gpu -x -q > $out/tests_all_fused2.log 2>&1; echo "rc=$?" >> $out/tests_all_fused2.log
tail -4 $out/tests_all_fused2.log
AFSK_FUSED=1 timeout 1500 python -m pytest tests -m gpu -x -q > $out/tests_all_fused1.log 2>&1; echo "rc=$?" >> $out/tests_all_fused1.log
tail -4 $out/tests_all_fused1.log
for wl in c2 c5 c4; do
  for f in 0 1 0 1 0 1; do
    AFSK_FUSED=$f timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 30 > $out/ab_${wl}_f${f}.json 2>> $out/ab.err
    python tools/benchline.py "$wl fused=$f" < $out/ab_${wl}_f${f}.json | tee -a $out/ab_summary.txt
  done
done
