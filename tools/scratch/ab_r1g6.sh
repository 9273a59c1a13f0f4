#!/bin/bash
O=gpurun_out/r1g; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_rx_one_mixed or random_sweep or every_alignment or fuzz or mixed_corpus or hamming or threshold_edge" 2>&1 | tail -3 | tee $O/pytest_pair.log
export AB_ROUNDS=4
for w in c3 w3000 w2000; do timeout 300 python tools/ab_demod.py $w "" 2>&1 | tee -a $O/ab6.txt; done
