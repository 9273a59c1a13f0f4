#!/bin/bash
O=gpurun_out/r1g; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_rx_one_mixed or random_sweep or every_alignment or fuzz or mixed_corpus or hamming" 2>&1 | tail -3 | tee $O/pytest_tie.log
export AB_ROUNDS=4
for w in c3 w3000 w2000; do
  for rep in 1 2; do
    echo -n "NEW  " | tee -a $O/ab4.txt; timeout 300 python tools/ab_demod.py $w "" 2>&1 | tee -a $O/ab4.txt
    echo -n "PREV " | tee -a $O/ab4.txt; AFSK_LIB_PATH=$PWD/afskmodem_b200/libafsk_b200_prev.so timeout 300 python tools/ab_demod.py $w "" 2>&1 | tee -a $O/ab4.txt
  done
done
