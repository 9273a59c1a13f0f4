#!/bin/bash
# same-box A/B of a side-by-side build (afskmodem_b200/libafsk_b200_alt.so) against the library of record
O=gpurun_out/alt; mkdir -p $O
export AB_ROUNDS=4
for w in "$@"; do
  for rep in 1 2; do
    echo -n "REC " | tee -a $O/ab.txt; timeout 300 python tools/ab_demod.py $w "" 2>&1 | tee -a $O/ab.txt
    echo -n "ALT " | tee -a $O/ab.txt; AFSK_LIB_PATH=$PWD/afskmodem_b200/libafsk_b200_alt.so timeout 300 python tools/ab_demod.py $w "" 2>&1 | tee -a $O/ab.txt
  done
done
