"""ctypes loader for the CPU oracle (``oracle/afsk_oracle.c``).

TEST INFRASTRUCTURE ONLY — the checker, never the product.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
``afskmodem_b200`` must not import this module (tests/test_no_oracle_in_product.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libafsk_oracle.so")

ST_OK, ST_NO_CLOCK, ST_NO_DATA = 0, 1, 2
ST_EXC_WAVELEN, ST_EXC_INDEX, ST_EXC_BAUD = -1, -2, -3

EXC_TEXT = {
    ST_EXC_WAVELEN: ("Exception", "Comparing two waveforms of different lengths."),
    ST_EXC_INDEX: ("IndexError", "list index out of range"),
    ST_EXC_BAUD: ("Exception", "Invalid baud rate."),
}


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("clock", C.c_int32), ("train_end", C.c_int64),
                ("nbits", C.c_int64), ("nbytes", C.c_int64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "afsk_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libafsk_oracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        i16p, u8p, i64p, i32p = (C.POINTER(C.c_int16), C.POINTER(C.c_uint8), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int32))
        L.afsk_oracle_space_tone.argtypes = [C.c_int, i16p]
        L.afsk_oracle_mark_tone.argtypes = [C.c_int, i16p]
        L.afsk_oracle_training_cycle.argtypes = [C.c_int, i16p]
        L.afsk_oracle_amplitude.argtypes = [i16p, C.c_int]
        L.afsk_oracle_ecc_decode.argtypes = [u8p, C.c_int64, u8p]
        L.afsk_oracle_ecc_decode.restype = C.c_int64
        L.afsk_oracle_ecc_encode.argtypes = [u8p, C.c_int64, u8p]
        L.afsk_oracle_ecc_encode.restype = C.c_int64
        L.afsk_oracle_rx_decode.argtypes = [i16p, C.c_int64, C.c_int, C.c_int, u8p, u8p, C.POINTER(Result)]
        L.afsk_oracle_rx_decode_batch.argtypes = [i16p, i64p, C.c_int, i32p, i32p, u8p, i64p,
                                                  C.POINTER(Result), C.c_int]
        L.afsk_oracle_listen_gate.argtypes = [i16p, C.c_int64, C.c_int, C.c_int, C.c_int64, i64p, i64p]
        L.afsk_oracle_tx_num_frames.argtypes = [C.c_int, C.c_int64, C.c_int64, u8p]
        L.afsk_oracle_tx_num_frames.restype = C.c_int64
        L.afsk_oracle_tx_frames.argtypes = [u8p, C.c_int64, C.c_int, C.c_int64, i16p, C.c_int64]
        L.afsk_oracle_tx_frames.restype = C.c_int64
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def tone(kind: str, baud: int) -> np.ndarray | None:
    """kind in {"space","mark","training"}; None ⇔ Exception("Invalid baud rate.")."""
    buf = np.zeros(2 * 48000 + 16, dtype=np.int16)
    n = getattr(lib(), {"space": "afsk_oracle_space_tone", "mark": "afsk_oracle_mark_tone",
                        "training": "afsk_oracle_training_cycle"}[kind])(int(baud), _p(buf, C.c_int16))
    return None if n < 0 else buf[:n].copy()


def amplitude(x: np.ndarray) -> int:
    x = np.ascontiguousarray(x, dtype=np.int16)
    return lib().afsk_oracle_amplitude(_p(x, C.c_int16), len(x))


def ecc_encode(bits: np.ndarray) -> np.ndarray:
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(len(bits) // 4 * 7 + 8, dtype=np.uint8)
    n = lib().afsk_oracle_ecc_encode(_p(bits, C.c_uint8), len(bits), _p(out, C.c_uint8))
    return out[:n].copy()


def ecc_decode(bits: np.ndarray) -> np.ndarray:
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(len(bits) // 7 * 4 + 8, dtype=np.uint8)
    n = lib().afsk_oracle_ecc_decode(_p(bits, C.c_uint8), len(bits), _p(out, C.c_uint8))
    return out[:n].copy()


def rx_decode(samples: np.ndarray, baud: int, amp_end: int = 14000, want_bits: bool = False) -> dict:
    """Restated Receiver(baud, _, amp_end).load → dict(status, clock, train_end, nbits, nbytes, data[, bits])."""
    x = np.ascontiguousarray(samples, dtype=np.int16)
    n = len(x)
    out = np.zeros(n // 56 + 16, dtype=np.uint8)
    bits = np.zeros(n // 4 + 16, dtype=np.uint8) if want_bits else None
    r = Result()
    lib().afsk_oracle_rx_decode(_p(x, C.c_int16), n, int(baud), int(amp_end), _p(out, C.c_uint8),
                                _p(bits, C.c_uint8) if want_bits else None, C.byref(r))
    d = {"status": r.status, "clock": r.clock, "train_end": r.train_end, "nbits": r.nbits,
         "nbytes": r.nbytes, "data": out[:r.nbytes].tobytes()}
    if want_bits:
        d["bits"] = bits[:r.nbits].copy()
    return d


def rx_decode_batch(samples: np.ndarray, offsets: np.ndarray, baud, amp_end, threads: int = 1):
    """Batch of captures (CSR offsets) → (list of bytes, Result array as numpy structured view)."""
    x = np.ascontiguousarray(samples, dtype=np.int16)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    B = len(off) - 1
    baud = np.ascontiguousarray(np.broadcast_to(np.asarray(baud, dtype=np.int32), (B,)))
    amp_end = np.ascontiguousarray(np.broadcast_to(np.asarray(amp_end, dtype=np.int32), (B,)))
    lens = np.diff(off)
    out_off = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(lens // 56 + 16, out=out_off[1:])
    out = np.zeros(int(out_off[-1]), dtype=np.uint8)
    res = (Result * B)()
    lib().afsk_oracle_rx_decode_batch(_p(x, C.c_int16), _p(off, C.c_int64), B, _p(baud, C.c_int32),
                                      _p(amp_end, C.c_int32), _p(out, C.c_uint8), _p(out_off, C.c_int64),
                                      res, int(threads))
    datas = [out[out_off[c]:out_off[c] + res[c].nbytes].tobytes() for c in range(B)]
    return datas, res


def listen_gate(stream: np.ndarray, amp_start: int, amp_end: int, timeout_frames: int):
    """→ (recorded?, start, end) of Receiver.__listen over a finite recording."""
    s = np.ascontiguousarray(stream, dtype=np.int16)
    a, b = C.c_int64(0), C.c_int64(0)
    ok = lib().afsk_oracle_listen_gate(_p(s, C.c_int16), len(s), int(amp_start), int(amp_end),
                                       int(timeout_frames), C.byref(a), C.byref(b))
    return bool(ok), a.value, b.value


def listen_gate_multi(stream: np.ndarray, amp_start: int, amp_end: int, timeout_frames: int, max_calls: int = 0):
    """→ [(recorded?, start, end), ...] of successive Receiver.receive() calls over one recording."""
    s = np.ascontiguousarray(stream, dtype=np.int16)
    max_calls = max_calls or max(1, len(s) // 2048)
    out = np.zeros((max_calls, 3), dtype=np.int64)
    L = lib()
    L.afsk_oracle_listen_gate_multi.restype = C.c_int64
    n = L.afsk_oracle_listen_gate_multi(_p(s, C.c_int16), C.c_int64(len(s)), int(amp_start), int(amp_end),
                                        C.c_int64(int(timeout_frames)), C.c_int64(max_calls), _p(out, C.c_int64))
    return [(bool(a), int(b), int(c)) for a, b, c in out[:n]]


def ts_cycles(baud: int, training_time: float) -> int:
    """Transmitter.__init__ afskmodem.py:438 (Python float arithmetic, truncation)."""
    return int(baud * training_time / 2)


def tx_frames(payload: bytes, baud: int, training_time: float = 0.5) -> np.ndarray | None:
    """Restated Transmitter(baud, training_time).save → int16 frames in the wav; None ⇔ invalid baud."""
    pl = np.frombuffer(bytes(payload), dtype=np.uint8).copy() if len(payload) else np.zeros(1, dtype=np.uint8)
    nb = len(payload)
    tsc = ts_cycles(baud, training_time)
    need = lib().afsk_oracle_tx_num_frames(int(baud), tsc, nb, _p(pl, C.c_uint8))
    if need < 0:
        return None
    out = np.zeros(need + 8, dtype=np.int16)
    n = lib().afsk_oracle_tx_frames(_p(pl, C.c_uint8), nb, int(baud), tsc, _p(out, C.c_int16), len(out))
    assert n >= 0
    return out[:n].copy()
