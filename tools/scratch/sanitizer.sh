#!/bin/bash
O=gpurun_out/final; mkdir -p $O
python tools/scratch/files_breakdown.py 2>&1 | tee $O/files_breakdown.txt
K="golden_rx_one_mixed or random_sweep or every_alignment or fuzz or gate or golden_tx or unequal or threshold_edge or empty_and_ragged or ranges_plan"
for t in memcheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $t --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/san_$t.log 2>&1
  echo "$t rc=$?" | tee -a $O/san_summary.txt
  grep -E "ERROR SUMMARY|passed|failed" $O/san_$t.log | tail -3 | tee -a $O/san_summary.txt
done
