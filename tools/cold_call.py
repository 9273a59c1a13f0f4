#!/usr/bin/env python
"""Cold-path probe run by bench.py in a FRESH process: how long the first call of the public API takes.

    python tools/cold_call.py <device> <dir-with-cap*.wav | one.wav> <nfiles | 0> <baud>

nfiles > 0: first and second ``Receiver.load_batch`` over cap00000.wav .. in the directory;
nfiles == 0: first and later ``Receiver.load`` of the one file.  CUDA initialisation (driver, context,
module load) is timed separately: it belongs to the process, not to the call.
Prints one JSON line.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    device, path, nf, baud = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    t0 = time.perf_counter()
    import afskmodem_b200 as A
    from afskmodem_b200 import _cabi
    t_import = (time.perf_counter() - t0) * 1e3
    A.LOG_LEVEL = 5
    t0 = time.perf_counter()
    _cabi.require_device(device)
    warm = _cabi.DeviceBuffer(device, 1 << 20)        # forces driver + context initialisation
    _cabi.stream_sync(device)
    t_init = (time.perf_counter() - t0) * 1e3
    rx = A.Receiver(baud, device=device)
    out = {"import_ms": t_import, "cuda_init_ms": t_init}
    if nf > 0:
        names = [os.path.join(path, f"cap{c:05d}.wav") for c in range(nf)]
        t0 = time.perf_counter()
        a = rx.load_batch(names, string=False, errors="return", log=False)
        out["first_load_batch_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        b = rx.load_batch(names, string=False, errors="return", log=False)
        out["second_load_batch_ms"] = (time.perf_counter() - t0) * 1e3
        out["files"] = nf
        out["same_results"] = [x if isinstance(x, bytes) else b"" for x in a] == [x if isinstance(x, bytes) else b"" for x in b]
    else:
        t0 = time.perf_counter()
        first = rx.load(path, True)
        out["first_load_ms"] = (time.perf_counter() - t0) * 1e3
        lat = []
        for _ in range(20):
            t0 = time.perf_counter()
            rx.load(path, True)
            lat.append((time.perf_counter() - t0) * 1e3)
        out["later_load_ms_median"] = sorted(lat)[len(lat) // 2]
        out["decoded"] = first if isinstance(first, str) else first.decode("utf-8", "replace")
    warm.close()
    rx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
