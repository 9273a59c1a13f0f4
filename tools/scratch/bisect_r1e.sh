#!/bin/bash
O=gpurun_out/r1e; mkdir -p $O
for v in "" "AFSK_CLOCK_V1=1" "AFSK_FRAME_V1=1" "AFSK_SMALL_V1=1"; do
  echo "=== variant [$v]" | tee -a $O/bisect.log
  env $v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden_rx_one_mixed or random_sweep or every_alignment or gate or fuzz" 2>&1 | grep -E "passed|failed|FAILED|Error" | tee -a $O/bisect.log
done
