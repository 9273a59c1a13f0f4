#!/bin/bash
O=gpurun_out/r1g; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"k_demod" -s 3 -c 1 -o $O/rx_c4 -f python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_demod" -s 3 -c 1 -o $O/rx_w600 -f python bench.py --workload w600 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_w600.log 2>&1
ls -la $O | tail -4
