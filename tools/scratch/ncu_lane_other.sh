#!/bin/bash
O=gpurun_out/final; mkdir -p $O
for w in w3000 w2000; do
  ncu --set full --clock-control none --import-source on -k regex:"k_demod" -s 3 -c 1 -o $O/rx_$w -f python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_$w.log 2>&1
done
ls -la $O | tail -3
