#!/bin/bash
out=gpurun_out/r2b2; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "fused" > $out/tests_fused.log 2>&1; echo "rc=$?" >> $out/tests_fused.log
tail -4 $out/tests_fused.log
for wl in c2 c3 c5; do
  for f in 0 1 0 1; do
    AFSK_FUSED=$f timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/ab_${wl}_f${f}.json 2>> $out/ab.err
    python tools/benchline.py "$wl fused=$f" < $out/ab_${wl}_f${f}.json | tee -a $out/ab_summary.txt
  done
done
