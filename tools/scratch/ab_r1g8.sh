#!/bin/bash
mkdir -p gpurun_out/r1g
export AB_ROUNDS=4
timeout 400 python tools/ab_demod.py c3 "AFSK_L2_HINT=0" "AFSK_L2_HINT=2" "AFSK_L2_HINT=4" 2>&1 | tee gpurun_out/r1g/ab8_c3.txt
