"""first-call vs steady-state cost of Receiver.load_batch for staging-ring geometries (fresh process each)"""
import json, os, subprocess, sys, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import afskmodem_b200 as A
A.LOG_LEVEL = 5
tmp = tempfile.mkdtemp(prefix="afsk_cold_", dir="/dev/shm")
try:
    rng = np.random.default_rng(1)
    nf = 1024
    pay = [rng.integers(0, 256, 1024, dtype=np.uint8).tobytes() for _ in range(nf)]
    A.Transmitter(1200).save_batch(pay, [os.path.join(tmp, f"cap{c:05d}.wav") for c in range(nf)])
    for mb, slots in ((32, 6), (16, 6), (16, 4), (8, 4), (8, 8), (4, 8)):
        env = dict(os.environ, AFSK_WAV_SLOT_MB=str(mb), AFSK_WAV_SLOTS=str(slots))
        best = None
        for _ in range(2):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cold_call.py"), "0", tmp, str(nf), "1200"],
                                 capture_output=True, text=True, env=env, timeout=300)
            r = json.loads(out.stdout.strip().splitlines()[-1])
            if best is None or r["first_load_batch_ms"] < best["first_load_batch_ms"]:
                best = r
        print(f"slot {mb} MB x {slots}: first {best['first_load_batch_ms']:.1f} ms, second {best['second_load_batch_ms']:.1f} ms, "
              f"ratio {best['first_load_batch_ms'] / best['second_load_batch_ms']:.2f}, cuda init {best['cuda_init_ms']:.0f} ms", flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
