#!/bin/bash
out=gpurun_out/r2o; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "sweep or alignment or fuzz or mixed or fused_clock_every or golden" > $out/tests.log 2>&1; echo "rc=$?" >> $out/tests.log; tail -3 $out/tests.log
timeout 300 python bench.py --workload w12000 --no-extra --no-e2e --no-cpu-baseline --steps 20 2>> $out/err.log | python tools/benchline.py "w12000 k_demod_shift<4,8>"
