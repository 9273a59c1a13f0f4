#!/bin/bash
# round-1e: parity of the new k_demod_small / k_clock / k_frame_warp / k_gate_amp, then same-box A/B
O=gpurun_out/r1e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
export AB_ROUNDS=4
ALLV1="AFSK_SMALL_V1=1,AFSK_CLOCK_V1=1,AFSK_FRAME_V1=1"
timeout 300 python tools/ab_demod.py c2 "" "AFSK_CLOCK_V1=1" "AFSK_FRAME_V1=1" "$ALLV1" 2>&1 | tee $O/ab_c2.txt
timeout 300 python tools/ab_demod.py c3 "" "AFSK_SMALL_V1=1" "AFSK_CLOCK_V1=1" "AFSK_FRAME_V1=1" "$ALLV1" 2>&1 | tee $O/ab_c3.txt
timeout 300 python tools/ab_demod.py c5 "" "$ALLV1" 2>&1 | tee $O/ab_c5.txt
timeout 300 python tools/ab_demod.py w3000 "" "AFSK_SMALL_V1=1" 2>&1 | tee $O/ab_w3000.txt
timeout 300 python tools/ab_demod.py w2000 "" "AFSK_SMALL_V1=1" 2>&1 | tee $O/ab_w2000.txt
