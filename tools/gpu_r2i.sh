#!/bin/bash
# 8 GPUs: e2e with the gate (tail overlap), then the full default line
out=gpurun_out/r2i_n8; mkdir -p $out
for pol in 4:2 2:1; do
  AFSK_H2D_GATE=$pol timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra --no-cpu-baseline --e2e-steps 5 > $out/bench_gate_${pol/:/_}.json 2> $out/bench_gate_${pol/:/_}.err
  python - <<P
import json
d=json.loads(open("$out/bench_gate_${pol/:/_}.json").read().strip().splitlines()[-1])
print("gate $pol: e2e", round(d["e2e"]["value"]), "Msamples/s  aggregate H2D", round(d["e2e"]["aggregate_h2d_gbs"],1), "GB/s  ms", round(d["e2e"]["ms_per_step"],1), " one-process:", {k:(round(v,1) if isinstance(v,float) else v) for k,v in (d.get("e2e_one_process") or {}).items() if k in ("value","ms","equal_to_single_device","error")})
P
done
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > $out/bench_n8.json 2> $out/bench_n8.err ) 2> $out/bench_n8.time
python tools/benchline.py "N=8 full" < $out/bench_n8.json; cat $out/bench_n8.time | head -2
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 20 --warmup 5 > $out/bench_n4.json 2> $out/bench_n4.err )
python tools/benchline.py "N=4 full" < $out/bench_n4.json
