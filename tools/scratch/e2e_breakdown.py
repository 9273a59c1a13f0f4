import time, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
import afskmodem_b200 as A
from afskmodem_b200 import _cabi
import bench
A.LOG_LEVEL = 5
B = bench.set_workload("c2", 0)
samples, offsets, spec = bench.build_batch_on_gpu(B, 0, 0)
total = int(offsets[-1])
pin = _cabi.PinnedArray((total + 64,), np.int16)
pin.array[:total] = samples[:total].cpu().numpy()
rx = A.Receiver(1200, 18000, 14000)
rx.decode_batch(pin.array, offsets)
s = rx._session(np.ascontiguousarray(offsets, dtype=np.int64), 0)
def t(f, n=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def up(): s.upload(pin.array); _cabi.stream_sync(0)
def run(): s.run(); _cabi.stream_sync(0)
print("upload ms", t(up), "GB/s", total * 2 / t(up) / 1e6)
print("run ms", t(run))
print("download ms", t(lambda: s.download()))
print("session lookup ms", t(lambda: rx._session(np.ascontiguousarray(offsets, dtype=np.int64), 0)))
print("full ms", t(lambda: rx.decode_batch(pin.array, offsets)))
# raw torch H2D for comparison
hp = torch.empty(total, dtype=torch.int16).pin_memory()
d = torch.empty(total, dtype=torch.int16, device='cuda')
print("torch pinned H2D ms", t(lambda: d.copy_(hp, non_blocking=True)))
# two streams, half each
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h = total // 2
def two():
    with torch.cuda.stream(s1): d[:h].copy_(hp[:h], non_blocking=True)
    with torch.cuda.stream(s2): d[h:].copy_(hp[h:], non_blocking=True)
print("two-stream H2D ms", t(two))
