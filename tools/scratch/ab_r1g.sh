#!/bin/bash
O=gpurun_out/r1g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu.log
export AB_ROUNDS=4
timeout 300 python tools/ab_demod.py c3 "" "AFSK_SMALL_LANE=0" 2>&1 | tee $O/ab_c3.txt
timeout 300 python tools/ab_demod.py w3000 "" "AFSK_SMALL_LANE=0" 2>&1 | tee $O/ab_w3000.txt
timeout 300 python tools/ab_demod.py w2000 "" "AFSK_SMALL_LANE=0" 2>&1 | tee $O/ab_w2000.txt
timeout 300 python tools/ab_demod.py c5 "" "AFSK_SMALL_LANE=0" 2>&1 | tee $O/ab_c5.txt
