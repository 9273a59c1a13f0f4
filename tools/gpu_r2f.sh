#!/bin/bash
out=gpurun_out/r2f; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "clock_kernel or empty_tx" > $out/tests.log 2>&1; echo "rc=$?" >> $out/tests.log; tail -3 $out/tests.log
AFSK_CLOCK_KERNEL=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/tests_ck2.log 2>&1; echo "rc=$?" >> $out/tests_ck2.log; tail -3 $out/tests_ck2.log
for wl in c3 c2 c5; do for k in 1 2 1 2; do
  AFSK_CLOCK_KERNEL=$k timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/b_${wl}_k$k.json 2>> $out/err.log
  python tools/benchline.py "$wl clock_kernel=$k" < $out/b_${wl}_k$k.json | tee -a $out/summary.txt
done; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 24 --csv --log-file $out/launches_c3_k2.csv env AFSK_CLOCK_KERNEL=2 python bench.py --workload c3 --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > /dev/null 2>&1
grep -E "k_clock|k_frame|k_demod" $out/launches_c3_k2.csv | tail -6 | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 24 --csv --log-file $out/launches_c3_k1.csv env AFSK_CLOCK_KERNEL=1 python bench.py --workload c3 --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > /dev/null 2>&1
grep -E "k_clock|k_frame|k_demod" $out/launches_c3_k1.csv | tail -6 | cut -c1-200
