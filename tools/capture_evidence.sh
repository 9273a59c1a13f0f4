#!/bin/bash
# round 2, final evidence of the final code: suite, bench lines, baud sweep, ncu launch lists and full captures, sanitizer
out=${1:-gpurun_out/evidence}; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q > $out/tests_all.log 2>&1; rc=$?; echo "rc=$rc" >> $out/tests_all.log; tail -3 $out/tests_all.log
[ $rc -eq 0 ] || exit 1
( time timeout 900 python bench.py > $out/bench_c2_n1.json 2> $out/bench_c2_n1.err ) 2> $out/bench.time; head -2 $out/bench.time
python tools/benchline.py "c2 default" < $out/bench_c2_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_c2_reference.json 2> $out/bench_ref.err
timeout 900 python tools/baud_sweep.py $out/baud_sweep.json > $out/baud_sweep.log 2>&1; tail -2 $out/baud_sweep.log | cut -c1-150
for wl in c2 c3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file $out/launches_bench_$wl.csv python bench.py --workload $wl --steps 5 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > /dev/null 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_demod|k_clock|k_frame" -s 9 -c 3 -o $out/ncu_rx_$wl -f python bench.py --workload $wl --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py $out/ncu_rx_$wl.ncu-rep > $out/ncu_rx_${wl}_summary.txt 2>&1
done
for wl in w480 w750; do
  timeout 300 ncu --set full --clock-control none -k regex:"k_demod_shift" -s 3 -c 1 -o $out/ncu_rx_$wl -f python bench.py --workload $wl --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py $out/ncu_rx_$wl.ncu-rep > $out/ncu_rx_${wl}_summary.txt 2>&1
done
# gpurun brings back at most 64 MiB: the summaries travel, of the reports only the c3 one
rm -f $out/ncu_rx_c2.ncu-rep $out/ncu_rx_w480.ncu-rep $out/ncu_rx_w750.ncu-rep
[ $(du -sm gpurun_out | cut -f1) -lt 55 ] || rm -f $out/ncu_rx_c3.ncu-rep
for tool in memcheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "fused_equals or retargeted or ring or every_framing or clock_kernel_variant_2 or sharded_more or random_sweep or unequal or empty_and_ragged or clock_q or framing_search" > $out/sanitizer_$tool.log 2>&1; echo "rc=$?" >> $out/sanitizer_$tool.log
  tail -4 $out/sanitizer_$tool.log
done
ls $out
