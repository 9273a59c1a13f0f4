"""Synthesis time by baud on one GPU: 4096 payloads of 1 KB, 3 warm-up + 5 timed afsk_tx_synth calls (CUDA events) —
the unequal-tone rates (k_tx_nibscan + k_synth_var) and 1200 baud (k_synth) for comparison.

    python tools/tx_time.py
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import afskmodem_b200 as A
rng = np.random.default_rng(8)
pay = [rng.integers(0, 256, 1024, dtype=np.uint8).tobytes() for _ in range(4096)]
for baud in (4800, 8000, 960, 1600, 24000, 1200):
    tx = A.TxSession(pay, baud, int(baud * 0.5 / 2), 0)
    tx.upload()
    for _ in range(3): tx.run()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    t0.record()
    for _ in range(5): tx.run(st)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    frames = int(tx.out_len.astype(np.int64).sum())
    print(f"tx {baud:6d} baud  {frames/1e6:8.1f} M frames  {ms:.4f} ms  {2*frames/ms/1e6:7.0f} GB/s written")
    tx.close()
