"""Runs the UNMODIFIED reference (``/root/reference/afskmodem.py``) as the pin for the oracle.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``tests/golden/make_golden.py`` and (never on the
GPU box, where ``/root/reference`` does not exist) ad-hoc validation may import this.  The
product package ``afskmodem_b200`` never does.

What it provides
  * ``available()``                 – is the reference mounted here?
  * ``ref_load(samples, baud, ...)`` – writes a wav, calls ``Receiver.load`` (afskmodem.py:420-430)
    with ``LOG_LEVEL = 0`` and scrapes the four stage integers from the debug log
    (afskmodem.py:338, 368, 380, 427) plus the exception type/message if one propagates.
  * ``ref_save(payload, baud, training_time)`` – ``Transmitter.save`` (afskmodem.py:481-484) → int16.
  * ``ref_receive(stream, baud, ...)`` – drives ``Receiver.receive`` (afskmodem.py:402-417) from a
    buffer through the stub pyaudio, for the listen gate (afskmodem.py:299-319).
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import re
import sys
import tempfile
import wave

import numpy as np

REF_PATH = "/root/reference/afskmodem.py"
_STUB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stub")
_mod = None


def available() -> bool:
    return os.path.exists(REF_PATH)


def module():
    """Import the reference under a private name with the stub pyaudio ahead on sys.path."""
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError("reference not mounted at " + REF_PATH)
        sys.path.insert(0, _STUB_DIR)
        try:
            sys.modules.pop("pyaudio", None)
            spec = importlib.util.spec_from_file_location("_afskmodem_reference", REF_PATH)
            _mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(_mod)
        finally:
            sys.path.remove(_STUB_DIR)
    return _mod


def write_wav(path: str, samples: np.ndarray) -> None:
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(48000)
        f.writeframes(np.ascontiguousarray(samples, dtype="<i2").tobytes())


def read_wav(path: str) -> np.ndarray:
    with wave.open(path, "rb") as f:
        return np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").copy()


_RE = {
    "clock": re.compile(r"Recovered clock\. \(frame (\d+)\)"),
    "train_end": re.compile(r"Training sequence terminated on frame (\d+)"),
    "nbits": re.compile(r"Decoded (\d+) bits"),
    "nbytes": re.compile(r"Decoded (\d+) bytes"),
}


def _scrape(log: str) -> dict:
    out = {}
    for k, rx in _RE.items():
        m = rx.search(log)
        out[k] = int(m.group(1)) if m else None
    out["no_clock"] = "Failed to recover clock" in log
    out["no_data"] = "No data." in log
    out["timed_out"] = "Timed out." in log
    return out


def _tmpdir() -> str:
    return "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()


def ref_load(samples: np.ndarray, baud: int, amp_start: int = 18000, amp_end: int = 14000,
             string: bool = False) -> dict:
    """Receiver(baud, amp_start, amp_end).load(wav, string) on the reference itself."""
    m = module()
    res: dict = {"exc": None, "ret": None, "ctor_exc": None}
    m.LOG_LEVEL = 0
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        try:
            r = m.Receiver(baud, amp_start, amp_end)
        except Exception as e:  # noqa: BLE001 - parity on type+message
            res["ctor_exc"] = (type(e).__name__, str(e))
            return res
        fd, path = tempfile.mkstemp(suffix=".wav", dir=_tmpdir())
        os.close(fd)
        try:
            write_wav(path, samples)
            try:
                res["ret"] = r.load(path, string)
            except Exception as e:  # noqa: BLE001
                res["exc"] = (type(e).__name__, str(e))
        finally:
            os.unlink(path)
    res.update(_scrape(buf.getvalue()))
    return res


def ref_save(payload, baud: int, training_time: float = 0.5) -> np.ndarray:
    """Transmitter(baud, training_time).save(payload, wav) → the int16 frames in the file."""
    m = module()
    m.LOG_LEVEL = 5
    t = m.Transmitter(baud, training_time)
    fd, path = tempfile.mkstemp(suffix=".wav", dir=_tmpdir())
    os.close(fd)
    try:
        t.save(payload, path)
        return read_wav(path)
    finally:
        os.unlink(path)


def ref_receive(stream: np.ndarray, baud: int, amp_start: int, amp_end: int, timeout: float) -> dict:
    """Receiver.receive(timeout, False) with 2048-frame reads served from ``stream``."""
    m = module()
    pa = sys.modules[m.pyaudio.__name__]
    m.LOG_LEVEL = 0
    pa.PyAudio.feed = np.ascontiguousarray(stream, dtype="<i2").tobytes()
    buf = io.StringIO()
    res: dict = {"exc": None, "ret": None}
    try:
        with contextlib.redirect_stdout(buf):
            r = m.Receiver(baud, amp_start, amp_end)
            try:
                res["ret"] = r.receive(timeout, False)
            except Exception as e:  # noqa: BLE001
                res["exc"] = (type(e).__name__, str(e))
        res["reads"] = pa.PyAudio.last_stream.reads
    finally:
        pa.PyAudio.feed = None
    res.update(_scrape(buf.getvalue()))
    return res


def ref_receive_many(stream: np.ndarray, baud: int, amp_start: int, amp_end: int, timeout: float,
                     max_calls: int = 64) -> list[dict]:
    """Successive ``Receiver.receive(timeout, False)`` calls of ONE Receiver over one recorded stream
    (the reference keeps its input stream open between calls, afskmodem.py:283): one dict per call
    until the feed is exhausted (the stub raises EOFError, which ends the list)."""
    m = module()
    pa = sys.modules[m.pyaudio.__name__]
    m.LOG_LEVEL = 0
    pa.PyAudio.feed = np.ascontiguousarray(stream, dtype="<i2").tobytes()
    calls: list[dict] = []
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            r = m.Receiver(baud, amp_start, amp_end)
        st = pa.PyAudio.last_stream
        for _ in range(max_calls):
            buf = io.StringIO()
            res: dict = {"exc": None, "ret": None, "reads_before": st.reads}
            with contextlib.redirect_stdout(buf):
                try:
                    res["ret"] = r.receive(timeout, False)
                except EOFError:
                    break
                except Exception as e:  # noqa: BLE001
                    res["exc"] = (type(e).__name__, str(e))
            res["reads_after"] = st.reads
            res.update(_scrape(buf.getvalue()))
            calls.append(res)
    finally:
        pa.PyAudio.feed = None
    return calls


def load_file_job(job) -> bytes:
    """(wav path, baud, amp_end) -> what the UNMODIFIED reference's ``Receiver.load(path, False)`` returns
    (b"" where it raises), with its log silenced (``LOG_LEVEL = 5``) as a user timing it would run it.
    Picklable for ``multiprocessing.Pool`` — the timing harness of ``bench.py --impl reference --python``."""
    path, baud, amp_end = job
    m = module()
    m.LOG_LEVEL = 5
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            r = m.Receiver(baud, 18000, amp_end)
            out = r.load(path, False)
        except Exception:  # noqa: BLE001
            return b""
    return out if isinstance(out, bytes) else out.encode("utf-8")
