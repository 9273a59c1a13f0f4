#!/bin/bash
# evidence capture of the final code of a round (run under gpurun, one GPU): new k_demod_small / k_clock / k_frame_warp
set -x
O=gpurun_out/final; mkdir -p $O
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_c2_reference.json 2>$O/ref.err
python bench.py > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err
for w in c3 c4 c5; do python bench.py --workload $w > $O/bench_${w}_n1.json 2> $O/bench_${w}.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40 --csv --log-file $O/launches_bench_c2.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40 --csv --log-file $O/launches_bench_c3.csv python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_demod|k_clock|k_frame" -s 9 -c 3 -o $O/rx_c2 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_demod|k_clock|k_frame" -s 9 -c 3 -o $O/rx_c3 -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_c3.log 2>&1
ls -la $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
K="golden_rx_one_mixed or random_sweep or every_alignment or fuzz or gate or pipelined"
for t in memcheck initcheck; do
  timeout 900 compute-sanitizer --tool $t --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/san_$t.log 2>&1
  echo "$t rc=$?" | tee -a $O/san_summary.txt
  grep -E "ERROR SUMMARY|passed|failed" $O/san_$t.log | tail -3 | tee -a $O/san_summary.txt
done
