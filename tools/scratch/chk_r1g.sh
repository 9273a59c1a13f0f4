#!/bin/bash
O=gpurun_out/r1g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu2.log
for w in c3 w3000 w2000; do python bench.py --workload $w --no-e2e --no-cpu-baseline > $O/bench_${w}.json 2>$O/bench_${w}.err; python tools/benchline.py $w < $O/bench_${w}.json; done
