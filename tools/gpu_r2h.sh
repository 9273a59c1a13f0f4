#!/bin/bash
out=gpurun_out/r2h; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "tx or unequal or synth or save_batch or cli" > $out/tests_tx.log 2>&1; echo "rc=$?" >> $out/tests_tx.log; tail -3 $out/tests_tx.log
python - <<'P'
import sys, json, torch
sys.path.insert(0, ".")
import bench, afskmodem_b200 as A
A.LOG_LEVEL = 5
bench.Ctx.local, bench.Ctx.dev = 0, torch.device("cuda", 0)
class Args: no_files = False
r = bench.run_tx(Args)
print(json.dumps({k: v for k, v in r.items() if k in ("k_synth", "k_synth_var_4800", "e2e_save_batch")}))
P
timeout 300 python bench.py --workload c5 --no-extra --no-e2e --no-cpu-baseline --steps 10 | python tools/benchline.py "c5"
