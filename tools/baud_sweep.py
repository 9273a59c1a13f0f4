#!/usr/bin/env python
"""Demodulator throughput at every decodable baud of interest, one process, one GPU.

    python tools/baud_sweep.py [out.json]

For each baud: ~2.4 G samples of 1 KB captures with the bench's AWGN mix (bench.py --workload w<baud>), 256
captures checked against the oracle, W warm-up + K timed decodes (CUDA events), the dominant kernel's own
time from the plan's event pairs.  Writes one JSON object {baud: {...}} (kept as profiles/r2_baud_sweep.json,
which bench.py quotes in its line and shard.py's cost model follows).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import afskmodem_b200 as A  # noqa: E402

BAUDS = [240, 300, 375, 400, 480, 500, 600, 750, 800, 1000, 1200, 1500, 2000, 2400, 3000, 4000, 6000, 12000]


class Args:
    warmup, parity_captures, no_cpu_baseline = 3, 256, True


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "baud_sweep.json")
    A.LOG_LEVEL = 5
    bench.Ctx.local, bench.Ctx.dev = 0, torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peak, _ = bench.hbm_peak()
    table = {}
    for baud in BAUDS:
        rec, st = bench.run_workload(Args, f"w{baud}", 0, 10)
        st[4].close()
        del st
        torch.cuda.empty_cache()
        r = rec["roofline"]
        table[str(baud)] = {"bit_frames": 48000 // baud, "demod_gbs": round(r["achieved"], 1), "frac_of_copy_peak": round(r["frac"], 3),
                            "frac_of_read_only_ceiling": round(r["frac_of_read_only_ceiling"] or 0, 3),
                            "step_msamples_s": round(rec["value"]), "share_of_step": round(r["share_of_step"], 3),
                            "captures": rec["captures_per_gpu"], "parity_checked": rec["parity_checked_vs_oracle"]}
        print(baud, table[str(baud)], flush=True)
    json.dump({"what": "k_demod* GB/s of int16 samples by baud, B200, device-resident, tools/baud_sweep.py", "hbm_copy_peak_gbs": peak,
               "table": table}, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
