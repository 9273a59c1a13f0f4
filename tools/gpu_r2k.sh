#!/bin/bash
out=gpurun_out/r2k; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "sweep or alignment or fuzz or mixed or fused_clock_every" > $out/tests.log 2>&1; echo "rc=$?" >> $out/tests.log; tail -3 $out/tests.log
for wl in w800 w480 w400 w240; do
  AFSK_NO_LONG_SHIFT=1 timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 2>> $out/err.log | python tools/benchline.py "$wl general kernel"
  timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 2>> $out/err.log | python tools/benchline.py "$wl k_demod_shift"
done
