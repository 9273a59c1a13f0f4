#!/bin/bash
# A/B of compile-time variants: the library of record vs libafsk_b200_imad.so (t-add in the FMA pipe), separate processes alternated
O=gpurun_out/r1e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_rx_one_mixed or random_sweep or every_alignment or fuzz or mixed_corpus or config2 or long_capture" 2>&1 | tail -3 | tee $O/pytest_frame2.log
AFSK_LIB_PATH=$PWD/afskmodem_b200/libafsk_b200_imad.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_rx_one_mixed or random_sweep or every_alignment or fuzz" 2>&1 | tail -3 | tee $O/pytest_imad.log
export AB_ROUNDS=4
for w in c2 c3 w2400 c4; do
  for rep in 1 2; do
    echo -n "A " | tee -a $O/ab3.txt; timeout 300 python tools/ab_demod.py $w "" 2>&1 | tee -a $O/ab3.txt
    echo -n "B " | tee -a $O/ab3.txt; AFSK_LIB_PATH=$PWD/afskmodem_b200/libafsk_b200_imad.so timeout 300 python tools/ab_demod.py $w "" 2>&1 | tee -a $O/ab3.txt
  done
done
