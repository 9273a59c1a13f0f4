#!/bin/bash
# round 2, call A: new tests, whole GPU suite, the default bench line, the reference arm
out=gpurun_out/r2a; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/smi.txt 2>&1
nproc >> $out/smi.txt; free -g >> $out/smi.txt
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > $out/tests_round2.log 2>&1; echo "rc=$?" >> $out/tests_round2.log
tail -5 $out/tests_round2.log
timeout 1500 python -m pytest tests -m gpu -q -x > $out/tests_all.log 2>&1; echo "rc=$?" >> $out/tests_all.log
tail -5 $out/tests_all.log
( time timeout 900 python bench.py > $out/bench_c2.json 2> $out/bench_c2.err ) 2> $out/bench_c2.time; echo "rc=$?" >> $out/bench_c2.err
tail -c 600 $out/bench_c2.err; cat $out/bench_c2.time
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err ) 2> $out/bench_ref.time
cat $out/bench_ref.time
