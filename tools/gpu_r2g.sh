#!/bin/bash
# round 2, final evidence: suite, bench lines, baud sweep, ncu launch lists and full captures, sanitizer
out=gpurun_out/r2g; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/tests_all.log 2>&1; echo "rc=$?" >> $out/tests_all.log; tail -3 $out/tests_all.log
( time timeout 900 python bench.py > $out/bench_c2_n1.json 2> $out/bench_c2_n1.err ) 2> $out/bench.time; cat $out/bench.time | head -2
python tools/benchline.py "c2 default" < $out/bench_c2_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_c2_reference.json 2> $out/bench_ref.err
timeout 900 python tools/baud_sweep.py $out/baud_sweep.json > $out/baud_sweep.log 2>&1; tail -3 $out/baud_sweep.log
for wl in c2 c3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file $out/launches_bench_$wl.csv python bench.py --workload $wl --steps 5 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_demod|k_clock|k_frame" -s 9 -c 3 -o $out/ncu_rx_$wl -f python bench.py --workload $wl --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > /dev/null 2>&1
done
ls -la $out
