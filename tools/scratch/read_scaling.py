"""Host ingest alone: how fast do the library's reader threads fill the pinned buffer from /dev/shm wav files,
with and without the overlapped H2D, for several thread counts?  (tuning aid for afsk_wav_load)"""
import os, sys, time, tempfile, shutil, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import afskmodem_b200 as A
from afskmodem_b200 import _cabi
from afskmodem_b200.modem import WavBatch
from oracle import oracle as O
rng = np.random.default_rng(0)
nf = 1024
fr = O.tx_frames(rng.integers(0, 256, 1024, dtype=np.uint8).tobytes(), 1200, 0.5)
d = tempfile.mkdtemp(prefix="afsk_rs_", dir="/dev/shm")
names = [os.path.join(d, f"c{i:05d}.wav") for i in range(nf)]
samples = np.tile(fr, nf); starts = np.arange(nf, dtype=np.int64) * len(fr); lens = np.full(nf, len(fr), np.int64)
A.modem.write_wav_batch(names, samples, starts, lens)
print("host cpus", os.cpu_count(), "files", nf, "bytes", samples.nbytes)
pinned = None
dbuf = _cabi.DeviceBuffer(0, samples.nbytes + 64)
for threads in (4, 8, 12, 16, 24, 32):
    for with_h2d in (False, True):
        best = 1e9
        for rep in range(4):
            wb = WavBatch(names, threads)
            t0 = time.perf_counter()
            wb.read(pinned, 0, dbuf.ptr if with_h2d else None); pinned = wb.pinned
            _cabi.stream_sync(0)
            best = min(best, time.perf_counter() - t0)
        print(f"threads {threads:2d} h2d {int(with_h2d)}: {best*1e3:6.1f} ms  {samples.nbytes/best/1e9:5.1f} GB/s")
# plain memcpy rate of the host for reference: pinned -> pinned, one thread
a = pinned.array[:samples.size]; b = _cabi.PinnedArray((samples.size,), np.int16)
t0 = time.perf_counter(); b.array[:] = a; t1 = time.perf_counter()
print(f"numpy memcpy 1 thread: {samples.nbytes/(t1-t0)/1e9:.1f} GB/s")
shutil.rmtree(d)
