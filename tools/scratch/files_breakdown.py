import os, sys, time, tempfile, shutil, numpy as np
sys.path.insert(0, '/root/repo')
import afskmodem_b200 as A
from afskmodem_b200 import _cabi
from afskmodem_b200.modem import WavBatch
from oracle import oracle as O
A.LOG_LEVEL = 5
rng = np.random.default_rng(0)
nf = 1024
fr = O.tx_frames(rng.integers(0, 256, 1024, dtype=np.uint8).tobytes(), 1200, 0.5)
d = tempfile.mkdtemp(prefix="afsk_fb_", dir="/dev/shm")
names = [os.path.join(d, f"c{i:05d}.wav") for i in range(nf)]
samples = np.tile(fr, nf); starts = np.arange(nf, dtype=np.int64) * len(fr); lens = np.full(nf, len(fr), np.int64)
t = time.perf_counter(); A.modem.write_wav_batch(names, samples, starts, lens); print("write ms", (time.perf_counter() - t) * 1e3)
rx = A.Receiver(1200)
for rep in range(3):
    t0 = time.perf_counter(); wb = WavBatch(names); t1 = time.perf_counter()
    s = rx._session(wb.offsets, 0)
    if s.d_samples is None:
        s.d_samples = _cabi.DeviceBuffer(0, (wb.total * 2 + 15) // 16 * 16 + 16)
    t2 = time.perf_counter()
    wb.read(getattr(rx, "_pinned", None), 0, s.d_samples.ptr); rx._pinned = wb.pinned
    _cabi.stream_sync(0); t3 = time.perf_counter()
    s.run(); b = s.download(); t4 = time.perf_counter()
    out = [rx.to_python(b, i, False, False) for i in range(nf)]; t5 = time.perf_counter()
    print(f"rep {rep}: probe {1e3*(t1-t0):.1f} session {1e3*(t2-t1):.1f} read+h2d {1e3*(t3-t2):.1f} decode+d2h {1e3*(t4-t3):.1f} python {1e3*(t5-t4):.1f} total {1e3*(t5-t0):.1f} ms -> {wb.total/ (t5-t0)/1e6:.0f} Msamples/s")
t0 = time.perf_counter(); got = rx.load_batch(names, string=False, errors="return", log=False); print("load_batch ms", (time.perf_counter()-t0)*1e3)
shutil.rmtree(d)
