#!/bin/bash
mkdir -p gpurun_out/r1g
export AB_ROUNDS=4
timeout 400 python tools/ab_demod.py c3 "" "AFSK_LANE_J=5" "AFSK_LANE_J=5,AFSK_DEMOD_STAGES=4" "AFSK_LANE_J=6" "AFSK_LANE_J=6,AFSK_DEMOD_STAGES=4" "AFSK_LANE_J=7" "AFSK_LANE_J=10" "AFSK_L2_HINT=2" 2>&1 | tee gpurun_out/r1g/ab3_c3.txt
