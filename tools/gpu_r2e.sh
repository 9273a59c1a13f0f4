#!/bin/bash
out=gpurun_out/r2e; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or alignment or fuzz or mixed or pipelined" > $out/tests.log 2>&1; echo "rc=$?" >> $out/tests.log
tail -3 $out/tests.log
run() { # workload tpw nsub
  if [ "$2" = "-" ]; then unset AFSK_DEMOD_TPW_LOG2; else export AFSK_DEMOD_TPW_LOG2=$2; fi
  if [ "$3" = "-" ]; then unset AFSK_DEMOD_NSUB; else export AFSK_DEMOD_NSUB=$3; fi
  timeout 300 python bench.py --workload $1 --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/b_$1_$2_$3.json 2>> $out/err.log
  python tools/benchline.py "$1 tpw_log2=$2 nsub=$3" < $out/b_$1_$2_$3.json | tee -a $out/summary.txt
}
run w1500 - -; run w1500 0 2; run w1500 0 1; run w1500 2 8
run w750 - -; run w750 1 2; run w750 1 1; run w750 3 8
run w375 - -; run w375 2 2; run w375 2 1
run c2 - -; run c2 - 2
run w600 - -; run w600 - 2
run c4 - -; run c4 - 2
