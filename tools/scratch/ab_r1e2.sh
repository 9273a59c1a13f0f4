#!/bin/bash
O=gpurun_out/r1e; mkdir -p $O
export AB_ROUNDS=4
timeout 300 python tools/ab_demod.py c2 "" "AFSK_DEMOD_CTAS=3" "AFSK_DEMOD_CTAS=3,AFSK_DEMOD_STAGES=2" "AFSK_DEMOD_STAGES=4" 2>&1 | tee $O/ab2_c2.txt
timeout 300 python tools/ab_demod.py c3 "" "AFSK_DEMOD_CTAS=3,AFSK_DEMOD_STAGES=2" "AFSK_L2_HINT=1" "AFSK_L2_HINT=1,AFSK_DEMOD_CTAS=3,AFSK_DEMOD_STAGES=2" 2>&1 | tee $O/ab2_c3.txt
timeout 300 python tools/ab_demod.py c4 "" "AFSK_DEMOD_CTAS=3" "AFSK_DEMOD_CTAS=3,AFSK_DEMOD_STAGES=2" 2>&1 | tee $O/ab2_c4.txt
timeout 300 python tools/ab_demod.py w2400 "" "AFSK_DEMOD_CTAS=3" 2>&1 | tee $O/ab2_w2400.txt
timeout 300 python tools/ab_demod.py w4000 "" "AFSK_DEMOD_CTAS=3" "AFSK_DEMOD_CTAS=3,AFSK_DEMOD_STAGES=2" 2>&1 | tee $O/ab2_w4000.txt
