"""Host ingest / egress (afsk_wav_probe / afsk_wav_load / afsk_wav_save) against CPython's wave
module, which is what the reference uses (afskmodem.py:213-217, 256-263).  No GPU needed: the
readers fill a plain numpy buffer here (pinned memory + overlapped H2D is covered by -m gpu)."""
import os
import struct
import wave

import numpy as np
import pytest

A = pytest.importorskip("afskmodem_b200")


def _ref_load(path):
    """SoundInput.loadFromFile + __convertFrames (afskmodem.py:201-205, 213-217)."""
    with wave.open(path, "rb") as f:
        raw = f.readframes(f.getnframes())
    return np.frombuffer(raw[:len(raw) // 2 * 2], dtype="<i2")


def _ref_save(path, frames):
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(48000)
        f.writeframes(np.asarray(frames, dtype="<i2").tobytes())


def _chunk(name, body):
    return name + struct.pack("<L", len(body)) + body + (b"\0" if len(body) & 1 else b"")


def _fmt(tag=1, nch=1, rate=48000, bits=16):
    sw = (bits + 7) // 8
    return _chunk(b"fmt ", struct.pack("<HHLLHH", tag, nch, rate, rate * nch * sw, nch * sw, bits))


def _riff(*chunks, size=None):
    body = b"WAVE" + b"".join(chunks)
    return b"RIFF" + struct.pack("<L", len(body) if size is None else size) + body


@pytest.fixture
def files(tmp_path):
    rng = np.random.default_rng(7)
    pcm = lambda n: rng.integers(-32768, 32768, n, dtype=np.int16).astype("<i2").tobytes()  # noqa: E731
    cases = {
        "plain": _riff(_fmt(), _chunk(b"data", pcm(5000))),
        "empty": _riff(_fmt(), _chunk(b"data", b"")),
        "one_frame": _riff(_fmt(), _chunk(b"data", pcm(1))),
        "list_before_data": _riff(_fmt(), _chunk(b"LIST", b"INFOxyz"), _chunk(b"data", pcm(777))),
        "junk_before_fmt": _riff(_chunk(b"JUNK", b"\x01\x02\x03"), _fmt(), _chunk(b"data", pcm(300))),
        "trailing_chunk": _riff(_fmt(), _chunk(b"data", pcm(300)), _chunk(b"LIST", b"abcd")),
        "stereo": _riff(_fmt(nch=2), _chunk(b"data", pcm(2 * 400))),
        "stereo_ragged": _riff(_fmt(nch=2), _chunk(b"data", pcm(2 * 400 + 1))),      # nframes = size // 4
        "eight_bit": _riff(_fmt(bits=8), _chunk(b"data", pcm(250) + b"\x7f")),       # odd byte count: last byte dropped by pairing
        "rate_8k": _riff(_fmt(rate=8000), _chunk(b"data", pcm(64))),
        "fmt_18_bytes": _riff(_chunk(b"fmt ", struct.pack("<HHLLHHH", 1, 1, 48000, 96000, 2, 16, 0)), _chunk(b"data", pcm(99))),
        "data_size_too_big": _riff(_fmt(), b"data" + struct.pack("<L", 100000) + pcm(1234)),      # truncated file
        "riff_size_small": _riff(_fmt(), _chunk(b"data", pcm(1000)), size=4 + 24 + 8 + 500),      # RIFF size cuts the data chunk
        "odd_data_chunk": _riff(_fmt(bits=8), _chunk(b"data", b"\x01\x02\x03"), _chunk(b"LIST", b"zz")),
    }
    bad = {
        "not_riff": b"RIFX" + b"\0" * 64,
        "not_wave": b"RIFF" + struct.pack("<L", 40) + b"AVI " + b"\0" * 36,
        "data_before_fmt": _riff(_chunk(b"data", pcm(10)), _fmt()),
        "no_data": _riff(_fmt()),
        "float_format": _riff(_fmt(tag=3, bits=32), _chunk(b"data", pcm(64))),
        "zero_channels": _riff(_fmt(nch=0), _chunk(b"data", pcm(64))),
        "short": b"RIF",
        "extensible": _riff(_chunk(b"fmt ", struct.pack("<HHLLHHHHL", 0xFFFE, 1, 48000, 96000, 2, 16, 22, 16, 4) +
                                   b"\x01\x00\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"), _chunk(b"data", pcm(200))),
    }
    paths = {}
    for name, blob in {**cases, **bad}.items():
        p = str(tmp_path / (name + ".wav"))
        open(p, "wb").write(blob)
        paths[name] = p
    paths["missing"] = str(tmp_path / "does_not_exist.wav")
    return paths, list(cases), list(bad) + ["missing"]


def test_reader_equals_wave_module(files):
    paths, good, bad = files
    names = good + bad
    wb = A.WavBatch([paths[n] for n in names], threads=4)
    buf = np.zeros(wb.total + 64, dtype=np.int16)
    wb.read(buf)
    for i, n in enumerate(names):
        try:
            want = _ref_load(paths[n])
        except Exception as e:  # noqa: BLE001
            assert i in wb.errors and type(wb.errors[i]) is type(e) and str(wb.errors[i]) == str(e), n
            continue
        assert i not in wb.errors, (n, wb.errors.get(i))
        assert np.array_equal(wb.frames(i), want), n
    # the plain PCM cases were read by the native threads, not by the fallback
    assert all(wb.status[names.index(n)] == 0 for n in good)


def test_writer_is_byte_identical_to_wave_module(tmp_path):
    rng = np.random.default_rng(8)
    lens = [0, 1, 2, 4800, 35680, 100001]
    starts = np.cumsum([0] + lens[:-1]).astype(np.int64) + 3
    samples = rng.integers(-32768, 32768, int(starts[-1] + lens[-1] + 5), dtype=np.int16)
    mine = [str(tmp_path / f"m{i}.wav") for i in range(len(lens))]
    A.write_wav_batch(mine, samples, starts, np.array(lens, dtype=np.int64), threads=3)
    for i, n in enumerate(lens):
        ref = str(tmp_path / f"r{i}.wav")
        _ref_save(ref, samples[starts[i]:starts[i] + n])
        assert open(mine[i], "rb").read() == open(ref, "rb").read(), n
    with pytest.raises(OSError):
        A.write_wav_batch([str(tmp_path / "no_such_dir" / "x.wav")], samples, starts[:1], np.array([10], dtype=np.int64))


def test_many_files_round_trip(tmp_path):
    rng = np.random.default_rng(9)
    n = 300
    lens = rng.integers(0, 20000, n).astype(np.int64)
    starts = np.zeros(n, dtype=np.int64)
    np.cumsum(lens[:-1], out=starts[1:])
    samples = rng.integers(-32768, 32768, int(lens.sum()), dtype=np.int16)
    paths = [str(tmp_path / f"c{i:04d}.wav") for i in range(n)]
    A.write_wav_batch(paths, samples, starts, lens)
    wb = A.WavBatch(paths)
    assert np.array_equal(wb.nsamples[:n], lens) and not wb.errors
    got = wb.read(np.zeros(wb.total + 64, dtype=np.int16))
    assert np.array_equal(got, samples)
    assert np.array_equal(A.read_wav_frames(paths[17]), samples[starts[17]:starts[17] + lens[17]])
