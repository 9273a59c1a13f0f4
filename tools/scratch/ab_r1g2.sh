#!/bin/bash
O=gpurun_out/r1g; mkdir -p $O
export AB_ROUNDS=4
timeout 400 python tools/ab_demod.py c3 "" "AFSK_L2_HINT=1" "AFSK_LANE_J=4" "AFSK_LANE_J=4,AFSK_DEMOD_STAGES=4" "AFSK_LANE_J=4,AFSK_DEMOD_STAGES=5" "AFSK_LANE_J=2,AFSK_DEMOD_STAGES=6" "AFSK_LANE_J=4,AFSK_L2_HINT=1" "AFSK_DEMOD_STAGES=2" 2>&1 | tee $O/ab2_c3.txt
ncu --set full --clock-control none --import-source on -k regex:"k_demod" -s 3 -c 1 -o $O/lane_c3 -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_lane_c3.log 2>&1
