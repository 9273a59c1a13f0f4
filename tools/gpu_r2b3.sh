#!/bin/bash
out=gpurun_out/r2b3; mkdir -p $out
for wl in c2 c3; do
  for fk in 0 2; do
    AFSK_FUSED=1 AFSK_FRAME_KERNEL=$fk timeout 300 python bench.py --workload $wl --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/ab_${wl}_fk${fk}.json 2>> $out/ab.err
    python tools/benchline.py "$wl fused=1 frame_kernel=$fk" < $out/ab_${wl}_fk${fk}.json | tee -a $out/ab_summary.txt
  done
done
AFSK_FUSED=1 timeout 300 python bench.py --workload c2 --captures 512 --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/ab_c2_512.json 2>> $out/ab.err
python tools/benchline.py "c2 512 captures fused=1" < $out/ab_c2_512.json
AFSK_FUSED=0 timeout 300 python bench.py --workload c2 --captures 512 --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/ab_c2_512_0.json 2>> $out/ab.err
python tools/benchline.py "c2 512 captures fused=0" < $out/ab_c2_512_0.json
