"""Stub ``pyaudio`` so the UNMODIFIED reference (/root/reference/afskmodem.py) imports
without an audio device.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

The reference touches only ``PyAudio().open(...)``, ``paInt16`` and the ``Stream`` name on
the file path (afskmodem.py:183-190, 229-236).  ``read`` is overridable so tests can drive
the live ``receive()`` gate (afskmodem.py:299-319) from a buffer.
"""

paInt16 = 8


class Stream:
    """Feeds ``read(n)`` from ``PyAudio.feed`` (bytes) when set, else raises."""

    def __init__(self, feed=None):
        self._feed = feed
        self._pos = 0
        self.reads = 0
        self.written = []

    def start_stream(self):
        pass

    def stop_stream(self):
        pass

    def close(self):
        pass

    def read(self, n):
        if self._feed is None:
            raise RuntimeError("stub pyaudio: no input feed configured")
        self.reads += 1
        out = self._feed[self._pos:self._pos + 2 * n]
        self._pos += 2 * n
        if len(out) < 2 * n:
            raise EOFError("stub pyaudio: feed exhausted")
        return out

    def write(self, data, *a, **kw):
        self.written.append(bytes(data))


class PyAudio:
    feed = None          # class-level: bytes fed to every input stream opened after it is set
    last_stream = None

    def open(self, **kw):
        s = Stream(PyAudio.feed if kw.get("input") else None)
        PyAudio.last_stream = s
        return s
