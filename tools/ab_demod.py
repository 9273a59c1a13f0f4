"""A/B timing of k_demod variants on ONE GPU in ONE process (boxes differ by several percent in
clocks under the power cap, so variants are only comparable inside one call).

    python tools/ab_demod.py c2 "AFSK_DEMOD_STAGES=3" "AFSK_DEMOD_STAGES=4" ...

Each argument after the workload is a comma-separated list of NAME=VALUE environment settings read
by libafsk_b200 at plan creation / decode time.  Variants are run round-robin, several rounds; the
median k_demod time per variant is printed.
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import afskmodem_b200 as A  # noqa: E402
import bench  # noqa: E402


def main():
    wl = sys.argv[1]
    variants = sys.argv[2:] or [""]
    B = bench.resolve_workload(wl)[2]
    corpus = bench.Corpus(wl, B)
    samples, offsets = bench.build_on_gpu(corpus, 0, B, 0)
    spec = corpus.spec(0, B)
    total = int(offsets[-1])
    stream = torch.cuda.current_stream().cuda_stream
    knobs = sorted({kv.split("=")[0] for v in variants for kv in v.split(",") if kv})

    def setenv(v):
        for k in knobs:
            os.environ.pop(k, None)
        for kv in v.split(","):
            if kv:
                k, val = kv.split("=")
                os.environ[k] = val

    sessions = []
    for v in variants:
        setenv(v)
        s = A.RxSession(offsets, spec["baud_rx"], spec["amp_end"], 0)
        s.bind(samples.data_ptr(), samples.numel())
        sessions.append(s)
    times = [[] for _ in variants]
    steps = [[] for _ in variants]
    rounds = int(os.environ.get("AB_ROUNDS", "6"))
    step_only = os.environ.get("AB_STEP_ONLY") == "1"     # no events inside the decode (they break dependent launches)
    for r in range(rounds + 1):
        for i, (v, s) in enumerate(zip(variants, sessions)):
            setenv(v)
            s.set_timing(not step_only)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                s.run(stream)
            e1.record()
            torch.cuda.synchronize()
            ms, n = s.demod_time() if not step_only else (float("nan"), 0)
            s.set_timing(False)
            if r:                       # round 0 is warm-up
                times[i].append(ms / 20)
                steps[i].append(e0.elapsed_time(e1) / 20)
    for v, t, st in zip(variants, times, steps):
        med = statistics.median(t)
        print(f"{wl} {v or '(default)':40s} demod {med:.4f} ms  {2 * total / med / 1e6:7.0f} GB/s   step {statistics.median(st):.4f} ms"
              f"   [min {min(t):.4f} max {max(t):.4f}]")


if __name__ == "__main__":
    main()
