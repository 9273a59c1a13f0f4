#!/bin/bash
out=gpurun_out/r2j; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "mixed or fuzz or golden_rx_one or pipelined or sharded or fused_equals" > $out/tests.log 2>&1; echo "rc=$?" >> $out/tests.log; tail -3 $out/tests.log
for f in 0 1 0 1 0 1; do
  AFSK_GROUP_STREAMS=$f timeout 300 python bench.py --workload c5 --no-extra --no-e2e --no-cpu-baseline --steps 30 > $out/c5_gs$f.json 2>> $out/err.log
  python tools/benchline.py "c5 group_streams=$f" < $out/c5_gs$f.json
done
