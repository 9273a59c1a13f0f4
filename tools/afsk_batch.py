#!/usr/bin/env python
"""Batch counterpart of the reference's rx-demo-file.py / tx-demo-file.py (one file per prompt there,
any number of files per call here; same Receiver / Transmitter arguments, same printed messages).

    python tools/afsk_batch.py rx [--baud 1200] [--amp-end 14000] [--bytes] a.wav b.wav ...
    python tools/afsk_batch.py tx [--baud 1200] [--training-time 0.5] --out-dir DIR "message 1" "message 2" ...
    python tools/afsk_batch.py tx --out-dir DIR --from-files payload1.bin payload2.bin

rx prints one line per file: the decoded transmission, "Could not decode." (rx-demo-file.py:9-10)
or the exception the reference's load() would have raised for that file.  With --stages the four
debug-log integers of the reference (clock frame, training end frame, coded bits, bytes) follow.
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    rx = sub.add_parser("rx")
    rx.add_argument("--baud", type=int, default=1200)
    rx.add_argument("--amp-start", type=int, default=18000)
    rx.add_argument("--amp-end", type=int, default=14000)
    rx.add_argument("--bytes", action="store_true", help="print payloads as bytes (load(..., string=False))")
    rx.add_argument("--stages", action="store_true", help="also print clock / training end / bits / bytes")
    rx.add_argument("--device", type=int, default=0)
    rx.add_argument("--log-level", type=int, default=5, help="afskmodem.LOG_LEVEL (0 = the reference's debug lines)")
    rx.add_argument("files", nargs="+")
    tx = sub.add_parser("tx")
    tx.add_argument("--baud", type=int, default=1200)
    tx.add_argument("--training-time", type=float, default=0.5)
    tx.add_argument("--out-dir", required=True)
    tx.add_argument("--from-files", action="store_true", help="arguments are files holding the payload bytes")
    tx.add_argument("--device", type=int, default=0)
    tx.add_argument("messages", nargs="+")
    args = ap.parse_args(argv)

    import afskmodem_b200 as A
    if args.cmd == "rx":
        A.LOG_LEVEL = args.log_level
        r = A.Receiver(args.baud, args.amp_start, args.amp_end, device=args.device)
        out = r.load_batch(args.files, string=not args.bytes, errors="return")
        batch = r._cache[1].download() if args.stages else None
        for i, (f, v) in enumerate(zip(args.files, out)):
            if isinstance(v, Exception):
                line = f"{type(v).__name__}: {v}"
            elif len(v) == 0:
                line = "Could not decode."
            else:
                line = repr(v) if args.bytes else v
            if batch is not None and int(batch.status[i]) >= 0 and int(batch.clock[i]) >= 0 and not isinstance(v, Exception):
                line += (f"    [clock {int(batch.clock[i])}, training end {int(batch.train_end[i])}, "
                         f"{int(batch.nbits[i])} bits, {int(batch.nbytes[i])} bytes]")
            print(f"{f}: {line}")
        return 0
    os.makedirs(args.out_dir, exist_ok=True)
    if args.from_files:
        payloads = [open(m, "rb").read() for m in args.messages]
        names = [os.path.join(args.out_dir, os.path.basename(m) + ".wav") for m in args.messages]
    else:
        payloads = args.messages
        names = [os.path.join(args.out_dir, f"msg{i:05d}.wav") for i in range(len(payloads))]
    print("Saving to file...")                                   # tx-demo-file.py:11
    A.Transmitter(args.baud, args.training_time, device=args.device).save_batch(payloads, names)
    for n in names:
        print(n)
    print("Done.")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
