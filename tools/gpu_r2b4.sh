#!/bin/bash
out=gpurun_out/r2b4; mkdir -p $out
for v in dbg_noframe dbg_nosignal b200; do
  AFSK_BENCH_NOPARITY=1 AFSK_LIB_PATH=/root/repo/afskmodem_b200/libafsk_$v.so AFSK_FUSED=1 timeout 300 python bench.py --workload c2 --no-extra --no-e2e --no-cpu-baseline --steps 20 > $out/c2_$v.json 2>> $out/ab.err
  python tools/benchline.py "c2 fused $v" < $out/c2_$v.json
done
tail -3 $out/ab.err
