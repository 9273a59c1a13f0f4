#!/bin/bash
O=gpurun_out/r1g; mkdir -p $O
AFSK_LANE_WARPS=16 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_rx_one_mixed or random_sweep or every_alignment or fuzz or mixed_corpus" 2>&1 | tail -3 | tee $O/pytest_w16.log
export AB_ROUNDS=4
for w in c3 w3000 w2000; do timeout 300 python tools/ab_demod.py $w "" "AFSK_LANE_WARPS=16" 2>&1 | tee -a $O/ab7.txt; done
