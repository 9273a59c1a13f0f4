"""GPU (-m gpu): the CUDA path, called through the C ABI (ctypes), against the committed golden
outputs of the real reference and against the CPU oracle on seeded inputs.  Bit-exact: every
quantity on this path is integer/byte work."""
import hashlib
import os

import numpy as np
import pytest

from conftest import gate_multi_stream, load_gate_golden, load_gate_multi_golden, load_rx_golden, load_tx_golden

pytestmark = pytest.mark.gpu

A = pytest.importorskip("afskmodem_b200")
from afskmodem_b200 import _cabi  # noqa: E402
from oracle import oracle as O  # noqa: E402

RX = load_rx_golden()
TX = load_tx_golden()


@pytest.fixture(autouse=True, scope="module")
def _quiet_and_loaded():
    A.LOG_LEVEL = 5
    _cabi.require_device(0)      # no CPU fallback: fail loudly if the GPU/library is missing
    yield
    A.LOG_LEVEL = 0


def _impair(fr, rng, lead=0, gain=1.0, sigma=0.0, cut=None):
    x = np.concatenate([np.zeros(lead, np.int16), fr]).astype(np.float64) * gain
    if sigma > 0:
        x = x + np.round(rng.normal(0.0, sigma, size=len(x)))
    x = np.clip(np.trunc(x), -32768, 32767).astype(np.int16)
    return x if cut is None else x[:cut]


def _check_against_oracle(session_batch, caps, bauds, thrs):
    for i, x in enumerate(caps):
        o = O.rx_decode(x, int(bauds[i]), int(thrs[i]))
        got = (int(session_batch.status[i]), int(session_batch.clock[i]), int(session_batch.train_end[i]),
               int(session_batch.nbits[i]), int(session_batch.nbytes[i]), session_batch.payload(i))
        want = (o["status"], o["clock"], o["train_end"], o["nbits"], o["nbytes"], o["data"])
        assert got == want, f"capture {i} baud {bauds[i]} n {len(x)}: {got[:5]} != {want[:5]}"


def test_golden_rx_one_mixed_batch():
    """All golden captures (every baud class, every failure mode) decoded in ONE batch."""
    metas = [m for m, _ in RX if not m["ctor_exc"]]
    caps = [x for m, x in RX if not m["ctor_exc"]]
    samples, offsets = A.modem._concat(caps)
    s = A.RxSession(offsets, [m["baud"] for m in metas], [m["amp_end"] for m in metas])
    s.upload(samples)
    s.run()
    b = s.download()
    for i, m in enumerate(metas):
        name = m["name"]
        st = int(b.status[i])
        if m["exc"]:
            assert st < 0 and list(O.EXC_TEXT[st]) == m["exc"], name
            continue
        assert st >= 0, name
        assert b.payload(i).hex() == m["data_hex"], name
        assert int(b.clock[i]) == (-1 if m["clock"] is None else m["clock"]), name
        assert int(b.train_end[i]) == (-1 if m["train_end"] is None else m["train_end"]), name
        assert int(b.nbits[i]) == (m["nbits"] or 0), name
        assert int(b.nbytes[i]) == (m["nbytes"] or 0), name
        assert (st == _cabi.ST_NO_CLOCK) == bool(m["no_clock"]), name
    s.close()


@pytest.mark.parametrize("meta,x", RX, ids=[m["name"] for m, _ in RX])
def test_golden_rx_load_semantics(meta, x, tmp_path):
    """Receiver(...).load on a wav: return type / exception rules as the reference."""
    if meta["ctor_exc"]:
        with pytest.raises(Exception, match="Invalid baud rate."):
            A.Receiver(meta["baud"], meta["amp_start"], meta["amp_end"])
        return
    r = A.Receiver(meta["baud"], meta["amp_start"], meta["amp_end"])
    path = str(tmp_path / "c.wav")
    A.write_wav_frames(path, x)
    if meta["exc"]:
        with pytest.raises(Exception) as ei:
            r.load(path, False)
        assert [type(ei.value).__name__, str(ei.value)] == meta["exc"]
        return
    assert r.load(path, False).hex() == meta["data_hex"]
    if meta["string_exc"]:
        with pytest.raises(UnicodeDecodeError):
            r.load(path, True)
    else:
        ret = r.read(path, True)
        assert type(ret).__name__ == meta["string_ret_type"]
        if meta["string_ret"] is not None:
            assert ret == meta["string_ret"]


def test_golden_tx():
    for meta, fr in TX:
        pl = bytes.fromhex(meta["payload_hex"])
        if meta["exc"]:
            with pytest.raises(Exception, match="Invalid baud rate."):
                A.Transmitter(meta["baud"], meta["training_time"])
            continue
        t = A.Transmitter(meta["baud"], meta["training_time"])
        mine = t.encode_batch([pl]).frames(0)
        assert len(mine) == meta["n"], meta["name"]
        assert hashlib.sha256(mine.astype("<i2").tobytes()).hexdigest() == meta["sha256"], meta["name"]
        if fr is not None:
            assert np.array_equal(mine, fr), meta["name"]


def test_tx_batch_matches_oracle_mixed():
    rng = np.random.default_rng(5)
    # 4800 / 960 / 1600 / 8000 / 24000: mark tone two frames shorter than the space tone
    bauds = [300, 600, 800, 1200, 2400, 4000, 6000, 12000, 100, 1000, 1500, 2000, 3000, 4800, 960, 1600, 8000,
             24000, 4800, 1200]
    pls = [rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8).tobytes() for _ in bauds]
    tts = [float(rng.choice([0.5, 0.1, 0.02, 0.0])) for _ in bauds]
    s = A.TxSession(pls, bauds, [O.ts_cycles(b, t) for b, t in zip(bauds, tts)])
    s.upload(); s.run()
    out = s.download()
    for i, b in enumerate(bauds):
        assert np.array_equal(out.frames(i), O.tx_frames(pls[i], b, tts[i])), b
    s.close()


def test_tx_unequal_tones_long_payloads():
    """Unequal-tone bauds with payloads long enough to cross many scan blocks and output chunks,
    all training times, one batch mixed with an equal-tone capture."""
    rng = np.random.default_rng(48)
    cases = [(4800, 5000, 0.5), (960, 700, 0.1), (8000, 3000, 0.0), (24000, 2000, 0.02), (1600, 1, 0.5),
             (4800, 0, 0.5), (1200, 300, 0.1), (4800, 257, 0.013),
             # more codewords in one 65,536-frame chunk than k_synth_var stages in shared memory (4096): global searches
             (24000, 9000, 0.02), (8000, 12000, 0.1),
             # all-ones / all-zeros nibbles: the shortest and the longest codewords (no frames at all at 24000 baud)
             (24000, 5000, 0.02), (8000, 6000, 0.0), (4800, 9000, 0.1), (960, 2500, 0.5)]
    pls = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for _, n, _ in cases]
    pls[10] = b"\xff" * 5000
    pls[11] = b"\xff" * 3000 + b"\x00" * 3000
    pls[12] = b"\x00" * 4500 + b"\xff" * 4500
    s = A.TxSession(pls, [b for b, _, _ in cases], [O.ts_cycles(b, t) for b, _, t in cases])
    s.upload(); s.run()
    out = s.download()
    for i, (b, n, t) in enumerate(cases):
        want = O.tx_frames(pls[i], b, t)
        assert len(out.frames(i)) == len(want), (b, n)
        assert np.array_equal(out.frames(i), want), (b, n)
    s.close()
    # the reference's own save() at 4800 baud decodes nowhere (load raises), but the wav must match
    assert np.array_equal(A.Transmitter(4800).encode_batch([b"Hello World!"]).frames(0), O.tx_frames(b"Hello World!", 4800, 0.5))


@pytest.mark.parametrize("baud", [300, 600, 800, 1200, 2400, 4000, 6000, 12000, 100, 24, 375, 1000, 2000, 3000, 1500, 500, 750, 480, 400, 240,
                                  250, 160, 125, 60, 30])      # the last five: more than 128 clock chains per capture
def test_random_sweep_vs_oracle(baud):
    """Seeded impairments per baud: lead silence (arbitrary alignment), gain, AWGN, truncation,
    thresholds — final bytes AND the four stage integers must equal the oracle's."""
    rng = np.random.default_rng(baud)
    caps, thrs = [], []
    n_cases = 24 if baud >= 300 else 6
    for k in range(n_cases):
        pl = rng.integers(0, 256, int(rng.integers(1, 48)), dtype=np.uint8).tobytes()
        tt = float(rng.choice([0.5, 0.1, 0.05])) if baud >= 300 else 2.0
        fr = O.tx_frames(pl, baud, tt)
        x = _impair(fr, rng, lead=int(rng.integers(0, 5000)) if k % 2 else 0,
                    gain=float(rng.choice([1.0, 0.7, 0.45, 0.3])),
                    sigma=float(rng.choice([0, 2000, 8000, 14000, 19000, 26000])),
                    cut=int(rng.integers(2000, len(fr))) if k % 5 == 4 else None)
        caps.append(x)
        thrs.append(int(rng.choice([14000, 11000, 8000])))
    samples, offsets = A.modem._concat(caps)       # arbitrary (unaligned) capture starts
    s = A.RxSession(offsets, baud, thrs)
    s.upload(samples); s.run()
    b = s.download()
    _check_against_oracle(b, caps, [baud] * len(caps), thrs)
    # stage-level: raw coded bits of one capture equal the oracle's
    o = O.rx_decode(caps[0], baud, thrs[0], want_bits=True)
    if o["status"] == 0:
        bits, _ = s.planes(0)
        k0 = (o["train_end"] - o["clock"]) // (48000 // baud)
        assert np.array_equal(bits[k0:k0 + o["nbits"]].astype(np.uint8), o["bits"])
    s.close()


@pytest.mark.parametrize("baud", [1200, 300, 600, 2400, 4000, 6000, 3000, 2000, 1500, 750, 375, 800, 12000])
def test_every_alignment_and_clock_offset(baud):
    """Capture start alignment (mod 8 samples) x clock offset (lead silence) — exercises every
    (e0) weight table / unrolled alignment body of every demodulator variant (k_demod merge and
    plain, k_demod_small, k_demod_shift) and the misaligned TMA copies.  Payloads are long enough
    that a capture spans several tiles, the last one partial."""
    rng = np.random.default_rng([3, baud])
    fr = O.tx_frames(rng.integers(0, 256, 700 if baud >= 2000 else 40, dtype=np.uint8).tobytes(), baud, 0.1)
    caps = []
    for lead in range(0, 48):
        caps.append(_impair(fr, rng, lead=lead, sigma=3000))
        caps.append(np.zeros(int(rng.integers(1, 8)), np.int16))      # shifts the next capture's alignment
    samples, offsets = A.modem._concat(caps)
    s = A.RxSession(offsets, baud, 14000)
    s.upload(samples); s.run()
    _check_against_oracle(s.download(), caps, [baud] * len(caps), [14000] * len(caps))
    # stage level: every decision / quiet flag of a few captures, not only the decoded span
    for i in (0, 2, 14, 30):
        o = O.rx_decode(caps[i], baud, 14000, want_bits=True)
        if o["status"] == 0:
            bits, _ = s.planes(i)
            k0 = (o["train_end"] - o["clock"]) // (48000 // baud)
            assert np.array_equal(bits[k0:k0 + o["nbits"]].astype(np.uint8), o["bits"]), (baud, i)
    s.close()


def test_threshold_edge_samples():
    """Samples sitting exactly on the hard-limiter edges (+-512, +-513) and at full scale."""
    rng = np.random.default_rng(11)
    vals = np.array([-32768, -32767, -514, -513, -512, -511, -1, 0, 1, 511, 512, 513, 514, 32766, 32767], np.int16)
    caps = [rng.choice(vals, size=int(n)).astype(np.int16) for n in (4096, 5000, 9999, 20000)]
    fr = O.tx_frames(b"edge", 1200, 0.1).astype(np.int32)
    caps.append(np.where(fr > 0, 513, -513).astype(np.int16))
    caps.append(np.where(fr > 0, 512, -512).astype(np.int16))
    for thr in (14000, 400, 0, 1, 70000, -5):
        b = A.Receiver(1200, 18000, thr).decode_batch(caps)
        _check_against_oracle(b, caps, [1200] * len(caps), [thr] * len(caps))
    # the same sample soup through the other demodulator variants (ties between the mark and space
    # correlations are common here: the floor-compare path)
    for baud in (2400, 4000, 6000, 3000, 300, 800):
        b = A.Receiver(baud, 18000, 400).decode_batch(caps)
        _check_against_oracle(b, caps, [baud] * len(caps), [400] * len(caps))


def test_empty_and_ragged_batches():
    assert len(A.Receiver(1200).decode_batch([])) == 0
    caps = [np.zeros(0, np.int16), np.zeros(1, np.int16), np.zeros(4095, np.int16), np.zeros(4096, np.int16),
            O.tx_frames(b"", 1200, 0.5), O.tx_frames(b"x" * 300, 1200, 0.02)]
    b = A.Receiver(1200).decode_batch(caps)
    _check_against_oracle(b, caps, [1200] * len(caps), [14000] * len(caps))


def test_listen_gate_golden():
    hello = dict((m["name"], x) for m, x in RX)["readme_hello_1200"]
    for g in load_gate_golden():
        s = np.concatenate([np.zeros(g["lead"], np.int16), (hello.astype(np.float64) * g["gain"]).astype(np.int16),
                            np.zeros(g["tail_zeros"], np.int16)])
        r = A.Receiver(1200, g["amp_start"], g["amp_end"])
        rec, a, b = r.listen_gate(s, g["timeout"])[0]
        assert (rec, a, b) == O.listen_gate(s, g["amp_start"], g["amp_end"], int(g["timeout"] * 48000)), g["name"]
        if g["exc"]:
            assert not rec
            continue
        assert rec == (not g["timed_out"]), g["name"]
        got = r.receive_recording(s, g["timeout"], False)
        assert got.hex() == g["ret_hex"], g["name"]
        if rec:
            assert b // 2048 == g["reads"]


def test_listen_gate_streams_at_every_alignment():
    """Several recorded streams in one call whose first frames fall on every 16-byte phase (the chunk
    amplitudes are summed from 128-bit vectors with the frames outside a chunk masked off), full-scale
    negative samples included (|-32768| = 32768 must not wrap)."""
    rng = np.random.default_rng(77)
    streams = []
    for i in range(24):
        n = int(rng.integers(3, 9)) * 2048 + int(rng.integers(0, 2048)) + (i % 8 == 0) * 1   # ragged tails, odd lengths
        level = int(rng.choice([300, 9000, 15000, 19000, 32768]))
        x = rng.integers(-level, level, n).astype(np.int32)
        x[: int(rng.integers(1, 3)) * 2048] //= 64                      # quiet start, so that the gate opens later
        x[int(rng.integers(0, n))] = -32768
        if i % 5 == 0:
            x[2048:4096] = -32768                                       # a whole chunk at full-scale negative
        streams.append(np.clip(x, -32768, 32767).astype(np.int16))
    streams.append(np.zeros(2047, np.int16))                            # no full chunk at all
    streams.append(np.zeros(0, np.int16))
    r = A.Receiver(1200, 6000, 4000)            # mean |x| of the noise is level / 2
    opened = 0
    for timeout in (0.0, 0.05, 1.0):
        got = r.listen_gate(streams, timeout)
        want = [O.listen_gate(s, 6000, 4000, int(timeout * 48000)) for s in streams]
        assert got == want, timeout
        opened += sum(1 for g in got if g[0])
    assert opened >= 10                          # the comparison is not vacuous


@pytest.mark.parametrize("g", load_gate_multi_golden(), ids=[g["name"] for g in load_gate_multi_golden()])
def test_receive_all_matches_successive_reference_receives(g):
    """Recorded stream with several transmissions: gate on the GPU, all recordings decoded as one
    batch of non-adjacent ranges of the uploaded stream; equals successive reference receive() calls."""
    s = gate_multi_stream(g, O.tx_frames)
    r = A.Receiver(g["baud"], g["amp_start"], g["amp_end"])
    calls = r.listen_gate_multi(s, g["timeout"])
    assert calls == O.listen_gate_multi(s, g["amp_start"], g["amp_end"], int(g["timeout"] * 48000))
    got = r.receive_all(s, g["timeout"], False, keep_timeouts=True)
    want = [bytes.fromhex(c["ret_hex"]) for c in g["calls"]]
    assert got[:len(want)] == want and len(got) <= len(want) + 1
    assert r.receive_all(s, g["timeout"], False) == [w for w, c in zip(want, g["calls"]) if not c["timed_out"]] + got[len(want):]
    if got:
        assert r.receive_recording(s, g["timeout"], False) == got[0]


def test_pipelined_decode_equals_single_plan():
    """decode_batch(pipeline=K): upload and kernels overlapped over K capture ranges, each with its own
    plan over the same device buffer — every result field and payload equals the single-plan decode,
    for K that does and does not divide the batch, with exception / too-short / empty captures mixed in."""
    rng = np.random.default_rng(2024)
    caps, bauds, thrs = [], [], []
    for i in range(37):
        baud = int(rng.choice([300, 1200, 2400, 4000, 6000, 4800]))
        fr = O.tx_frames(rng.integers(0, 256, int(rng.integers(1, 60)), dtype=np.uint8).tobytes(),
                         6000 if baud == 4800 else baud)
        x = _impair(fr, rng, lead=int(rng.integers(0, 3000)), sigma=float(rng.choice([0, 3000, 9000])))
        if i % 9 == 4:
            x = x[:int(rng.integers(0, 4096))]                      # too short for clock recovery (or empty)
        caps.append(x); bauds.append(baud); thrs.append(int(rng.choice([14000, 9000])))
    samples, offsets = A.modem._concat(caps)
    rx = A.Receiver(1200)
    one = rx.decode_batch(samples, offsets, baud_rate=bauds, amp_end_threshold=thrs, pipeline=1)
    for k in (2, 3, 8, 37, 64):
        many = rx.decode_batch(samples, offsets, baud_rate=bauds, amp_end_threshold=thrs, pipeline=k)
        assert isinstance(rx._cache[1], A.modem.PipelinedRxSession) and len(rx._cache[1].sessions) <= k
        assert np.array_equal(many.results, one.results), k
        assert many.payloads() == one.payloads(), k
    for i, x in enumerate(caps):                                    # and the single plan equals the oracle
        if bauds[i] == 4800:
            assert int(one.status[i]) == (_cabi.ST_EXC_WAVELEN if len(x) >= 4096 else _cabi.ST_NO_CLOCK)
            continue
        o = O.rx_decode(x, bauds[i], thrs[i])
        assert (int(one.status[i]), int(one.clock[i]), int(one.nbits[i]), one.payload(i)) == \
               (o["status"], o["clock"], o["nbits"], o["data"]), i
    # pinned host memory, the automatic choice: a batch of 64 MB and more is pipelined
    big = [np.zeros(5_000_000, np.int16) for _ in range(8)]
    big[3][1000:1000 + len(caps[0])] = caps[0]
    bs, bo = A.modem._concat(big)
    auto = rx.decode_batch(bs, bo, baud_rate=[bauds[0]] * 8, amp_end_threshold=[thrs[0]] * 8)
    assert isinstance(rx._cache[1], A.modem.PipelinedRxSession)
    ref = rx.decode_batch(bs, bo, baud_rate=[bauds[0]] * 8, amp_end_threshold=[thrs[0]] * 8, pipeline=1)
    assert np.array_equal(auto.results, ref.results) and auto.payloads() == ref.payloads()
    # batches a caller keeps are never overwritten by later calls on the same receiver
    big[5][2000:2000 + len(caps[1])] = caps[1]
    bs2, _ = A.modem._concat(big)
    kw = {"baud_rate": [bauds[0]] * 8, "amp_end_threshold": [thrs[0]] * 8}
    kept = [rx.decode_batch(bs if i % 2 == 0 else bs2, bo, **kw) for i in range(6)]
    assert isinstance(rx._cache[1], A.modem.PipelinedRxSession)
    for i, b in enumerate(kept):
        want = kept[i % 2]
        assert np.array_equal(b.results, want.results) and b.payloads() == want.payloads(), i
    assert not np.array_equal(kept[0].results, kept[1].results)
    rx.close()
    assert kept[0].payload(3) == auto.payload(3)                    # still readable after the receiver is closed


def test_ranges_plan_equals_adjacent_plan():
    """afsk_rx_plan_create_ranges over non-adjacent, overlapping and out-of-order ranges of one buffer."""
    rng = np.random.default_rng(21)
    caps = [_impair(O.tx_frames(bytes(rng.integers(0, 256, 20, dtype=np.uint8)), 1200, 0.05), rng, lead=int(rng.integers(0, 999)),
                    sigma=5000) for _ in range(6)]
    samples, offsets = A.modem._concat(caps)
    starts = np.array([offsets[4], offsets[0], offsets[2] + 3, offsets[2], 0, offsets[5]], dtype=np.int64)
    lens = np.array([len(caps[4]), len(caps[0]), len(caps[2]) - 3, len(caps[2]) + 4500, 0, len(caps[5]) - 77], dtype=np.int64)
    s = A.RxSession(starts, 1200, 14000, lengths=lens)
    s.upload(samples); s.run()
    b = s.download()
    views = [samples[int(a):int(a + n)] for a, n in zip(starts, lens)]
    _check_against_oracle(b, views, [1200] * len(views), [14000] * len(views))
    s.close()


def test_mixed_corpus_per_capture_settings():
    """BASELINE config #5 in miniature: one batch mixing bauds (including 4800, where load raises, and
    9600, where the constructor does), per-capture amp_end thresholds with matching gains, training
    times, payload sizes and noise — through Receiver.decode_batch's per-capture overrides."""
    rng = np.random.default_rng(55)
    B = 96
    baud_rx = rng.choice([300, 600, 1200, 2400, 4000, 6000, 4800, 9600], size=B, p=[.15, .15, .15, .15, .15, .15, .05, .05])
    caps, thr = [], []
    for c in range(B):
        btx = 6000 if baud_rx[c] == 9600 else int(baud_rx[c])
        pl = rng.integers(0, 256, int(np.exp(rng.uniform(np.log(16), np.log(600)))), dtype=np.uint8).tobytes()
        pair = int(rng.integers(0, 3))
        thr.append([14000, 11000, 8000][pair])
        gain = float(rng.choice([[1.0, 0.7], [1.0, 0.7, 0.45], [1.0, 0.7, 0.45, 0.3]][pair]))
        fr = O.tx_frames(pl, btx, float(rng.choice([0.5, 1.5, 0.1, 0.02])))
        caps.append(_impair(fr, rng, lead=int(rng.integers(0, 4000)) if c % 4 == 0 else 0, gain=gain,
                            sigma=float(rng.choice([0, 2000, 8000, 14000, 19000, 26000]))))
    r = A.Receiver(1200)
    b = r.decode_batch(caps, baud_rate=baud_rx, amp_end_threshold=thr)
    _check_against_oracle(b, caps, baud_rx, thr)
    assert ((b.status < 0) == ((baud_rx == 9600) | (baud_rx == 4800))).all()
    i9600, i4800 = int(np.argmax(baud_rx == 9600)), int(np.argmax(baud_rx == 4800))
    with pytest.raises(Exception, match="Invalid baud rate."):
        r.to_python(b, i9600)
    with pytest.raises(Exception, match="Comparing two waveforms of different lengths."):
        r.to_python(b, i4800)


def test_save_batch_load_batch_files(tmp_path):
    """Transmitter.save_batch -> wav files -> Receiver.load_batch (native threaded ingest into pinned
    memory, spans streamed to the GPU while later files are read): same values as per-file load()."""
    rng = np.random.default_rng(77)
    n = 40
    msgs = ["Hello World!", "Héellóo World!"] + [bytes(rng.integers(32, 127, int(rng.integers(1, 400)), dtype=np.uint8)).decode()
                                                   for _ in range(n - 2)]
    paths = [str(tmp_path / f"m{i}.wav") for i in range(n)]
    t = A.Transmitter(1200, 0.1)
    t.save_batch(msgs, paths)
    for i in (0, 1, 7):
        assert np.array_equal(A.read_wav_frames(paths[i]), O.tx_frames(msgs[i].encode("utf-8"), 1200, 0.1))
    r = A.Receiver(1200)
    assert r.load_batch(paths) == msgs
    assert r.load_batch(paths, string=False) == [m.encode("utf-8") for m in msgs]       # cached plan + pinned buffer
    assert [r.load(p) for p in paths[:3]] == msgs[:3]
    # a file the wave module rejects, a missing file, a too-short capture and an undecodable one
    bad = str(tmp_path / "bad.wav")
    open(bad, "wb").write(b"RIFF\x04\x00\x00\x00WAVX")
    short = str(tmp_path / "short.wav")
    A.write_wav_frames(short, np.zeros(100, np.int16))
    noise = str(tmp_path / "noise.wav")
    A.write_wav_frames(noise, rng.integers(-32768, 32768, 60000, dtype=np.int16))
    mixed = [paths[0], bad, str(tmp_path / "missing.wav"), short, noise, paths[1]]
    got = r.load_batch(mixed, string=False, errors="return")
    assert got[0] == msgs[0].encode() and got[5] == msgs[1].encode("utf-8") and got[3] == b""
    assert type(got[1]).__name__ == "Error" and isinstance(got[2], FileNotFoundError)
    assert got[4] == O.rx_decode(A.read_wav_frames(noise), 1200, 14000)["data"]
    with pytest.raises(Exception):
        r.load_batch(mixed)
    r.close()


def test_full_size_roundtrip_property():
    """BASELINE config #2 shape (1200 baud, 1 KB payloads, 602,400 samples each) at B=256:
    GPU synth -> AWGN (sigma=8000) -> GPU decode returns every payload; the clean batch returns
    clock 0 / 14,336 bits everywhere (size-independent round-trip property)."""
    rng = np.random.default_rng(2)
    B = 256
    pls = [rng.integers(0, 256, 1024, dtype=np.uint8).tobytes() for _ in range(B)]
    tx = A.Transmitter(1200).encode_batch(pls)
    assert int(tx.out_len[0]) == 602400
    rx = A.Receiver(1200)
    s = A.RxSession(tx.out_off, 1200, 14000)
    s.upload(tx.samples); s.run()
    b = s.download()
    assert (b.status == 0).all() and (b.clock == 0).all() and (b.nbits == 14336).all()
    assert b.payloads() == pls
    noisy = np.clip(tx.samples.astype(np.int32) + np.round(rng.normal(0, 8000, len(tx.samples))).astype(np.int32),
                    -32768, 32767).astype(np.int16)
    s.upload(noisy); s.run()
    b = s.download()
    assert b.payloads() == pls and (b.nbits == 14336).all()
    # spot-check the noisy decode of a few captures against the oracle, stage integers included
    for i in (0, 100, 255):
        x = noisy[int(tx.out_off[i]):int(tx.out_off[i]) + 602400]
        o = O.rx_decode(x, 1200, 14000)
        assert (int(b.clock[i]), int(b.train_end[i]), int(b.nbits[i]), b.payload(i)) == \
               (o["clock"], o["train_end"], o["nbits"], o["data"])
    s.close()
    del rx


def test_long_capture_300_baud():
    """BASELINE config #4 shape, shortened: 300 baud (160 samples/bit, 4 threads per window), 8 KB payload."""
    rng = np.random.default_rng(4)
    pl = rng.integers(0, 256, 8192, dtype=np.uint8).tobytes()
    fr = A.Transmitter(300).encode_batch([pl]).frames(0)
    assert np.array_equal(fr, O.tx_frames(pl, 300))
    x = _impair(fr, rng, lead=1234, sigma=8000)
    b = A.Receiver(300).decode_batch([x, fr])
    _check_against_oracle(b, [x, fr], [300, 300], [14000, 14000])
    assert b.payload(1) == pl


def test_abi_rejects_bad_arguments():
    L = _cabi.lib()
    assert L.afsk_rx_decode(None, None, None, None, None) == _cabi.AFSK_E_ARG
    off = np.array([0, 10], dtype=np.int64)
    s = A.RxSession(off, 1200, 14000)
    d = _cabi.DeviceBuffer(0, 64)
    import ctypes as C
    rc = L.afsk_rx_decode(s.plan, C.c_void_p(d.ptr + 2), C.c_void_p(s.d_out.ptr), C.c_void_p(s.d_res.ptr), None)
    assert rc == _cabi.AFSK_E_ARG and b"aligned" in L.afsk_last_error()
    s.close(); d.close()


def test_batch_cli_round_trip(tmp_path, capsys):
    """tools/afsk_batch.py (batch counterpart of the reference's tx-demo-file.py / rx-demo-file.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("afsk_batch", os.path.join(os.path.dirname(__file__), "..", "tools", "afsk_batch.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    msgs = ["Hello World!", "Héellóo World!", "third"]
    assert cli.main(["tx", "--out-dir", str(tmp_path), *msgs]) == 0
    files = sorted(str(p) for p in tmp_path.glob("*.wav"))
    assert len(files) == 3
    # the files are what the reference's save() writes
    for f, m in zip(files, msgs):
        assert np.array_equal(A.modem.read_wav_frames(f), O.tx_frames(m.encode(), 1200, 0.5))
    (tmp_path / "short.wav").write_bytes(open(files[0], "rb").read()[:44 + 2 * 1000])      # < 4096 frames
    capsys.readouterr()
    assert cli.main(["rx", "--stages", *files, str(tmp_path / "short.wav")]) == 0
    out = capsys.readouterr().out.splitlines()
    assert out[0].startswith(f"{files[0]}: Hello World!    [clock 0, training end 24160, 168 bits, 12 bytes]")
    assert "Héellóo World!" in out[1] and "third" in out[2]
    assert out[3].endswith("Could not decode.")
    A.LOG_LEVEL = 5


def test_host_buffer_entry_points_like_the_integration_stub():
    """afsk_rx_decode_host / afsk_tx_synth_host / afsk_tx_num_samples / afsk_rx_out_capacity called with
    plain ctypes arrays exactly as the stub in INTEGRATION.md does (no plan, no session)."""
    import ctypes as C
    L = _cabi.lib()
    rng = np.random.default_rng(77)
    # --- Transmitter.save stub: one payload per call
    for baud, tt, msg in [(1200, 0.5, b"Hello World!"), (300, 0.1, b""), (6000, 0.5, bytes(range(256))), (4800, 0.5, b"uneven tones")]:
        ts = int(baud * tt / 2)
        pay = (C.c_uint8 * max(len(msg), 1)).from_buffer_copy(msg or b"\0")
        n = L.afsk_tx_num_samples(baud, ts, len(msg), pay if baud == 4800 else None)
        want = O.tx_frames(msg, baud, tt)
        assert n == len(want), (baud, n, len(want))
        out = np.zeros(n + 8, dtype=np.int16)
        rc = L.afsk_tx_synth_host(0, pay, (C.c_int64 * 2)(0, len(msg)), 1, (C.c_int32 * 1)(baud), (C.c_int64 * 1)(ts),
                                  out.ctypes.data_as(C.POINTER(C.c_int16)), (C.c_int64 * 2)(0, n + 8))
        assert rc == 0, L.afsk_last_error()
        assert np.array_equal(out[:n], want), baud
    # --- Receiver.load stub: one capture per call, and a small batch whose first offset is not 0
    caps = [_impair(O.tx_frames(b"Hello World!", 1200, 0.5), rng, lead=123, sigma=5000),
            _impair(O.tx_frames(rng.integers(0, 256, 200, dtype=np.uint8).tobytes(), 1200, 0.1), rng, lead=7, sigma=9000),
            np.zeros(100, np.int16), rng.integers(-32768, 32767, 9000).astype(np.int16)]
    for x in caps:
        cap = L.afsk_rx_out_capacity(len(x), 1200)
        out = (C.c_uint8 * cap)()
        res = _cabi.RxResult()
        rc = L.afsk_rx_decode_host(0, x.ctypes.data_as(C.POINTER(C.c_int16)), (C.c_int64 * 2)(0, len(x)), 1,
                                   (C.c_int32 * 1)(1200), (C.c_int32 * 1)(14000), out, (C.c_int64 * 2)(0, cap), C.byref(res))
        assert rc == 0, L.afsk_last_error()
        o = O.rx_decode(x, 1200, 14000)
        assert (res.status, res.clock, res.train_end, res.nbits, res.nbytes) == (o["status"], o["clock"], o["train_end"], o["nbits"], o["nbytes"])
        assert bytes(out[:res.nbytes]) == o["data"]
    pad = 37                                         # captures start at sample 37 of the host buffer
    samples = np.concatenate([np.zeros(pad, np.int16)] + caps)
    off = np.cumsum([pad] + [len(c) for c in caps]).astype(np.int64)
    B = len(caps)
    caps_b = [int(L.afsk_rx_out_capacity(len(c), 1200)) for c in caps]
    out_off = np.concatenate([[0], np.cumsum(caps_b)]).astype(np.int64)
    out = np.zeros(int(out_off[-1]), np.uint8)
    res = (_cabi.RxResult * B)()
    rc = L.afsk_rx_decode_host(0, samples.ctypes.data_as(C.POINTER(C.c_int16)), off.ctypes.data_as(C.POINTER(C.c_int64)), B,
                               (C.c_int32 * B)(*[1200] * B), (C.c_int32 * B)(*[14000] * B),
                               out.ctypes.data_as(C.POINTER(C.c_uint8)), out_off.ctypes.data_as(C.POINTER(C.c_int64)), res)
    assert rc == 0, L.afsk_last_error()
    for i, x in enumerate(caps):
        o = O.rx_decode(x, 1200, 14000)
        assert (res[i].status, res[i].clock, res[i].nbits) == (o["status"], o["clock"], o["nbits"]), i
        assert out[out_off[i]:out_off[i] + res[i].nbytes].tobytes() == o["data"], i


def test_fuzz_arbitrary_signals_mixed_batch():
    """Not modem signals at all: noise of several distributions, DC offsets, square waves at wrong rates,
    sparse impulses, saturated runs — random lengths, every demodulator variant in ONE mixed batch.  The
    reference decodes such input to *something* (or nothing); the GPU must agree bit for bit."""
    rng = np.random.default_rng(2024)
    bauds_pool = [300, 600, 1200, 2400, 4000, 6000, 3000, 2000, 1500, 800, 750, 480, 375, 12000]
    caps, bauds, thrs = [], [], []
    for k in range(140):
        n = int(rng.choice([0, 1, 4095, 4096, 4097, int(rng.integers(4098, 40000))], p=[.02, .02, .03, .05, .05, .83]))
        kind = k % 7
        if kind == 0:
            x = rng.integers(-32768, 32768, n)
        elif kind == 1:
            x = np.round(rng.normal(0, float(rng.choice([300, 520, 5000, 20000])), n))
        elif kind == 2:
            x = np.round(rng.normal(float(rng.choice([-20000, 600, 15000])), 3000, n))           # DC offset
        elif kind == 3:
            per = int(rng.integers(3, 90))
            x = np.where((np.arange(n) // per) % 2 == 0, 30000, -30000) + rng.integers(-500, 500, n)   # wrong-rate square
        elif kind == 4:
            x = np.zeros(n); idx = rng.integers(0, max(n, 1), n // 50); x[idx] = rng.choice([-32768, 32767], len(idx))
        elif kind == 5:
            x = np.cumsum(rng.integers(-900, 901, n))                                               # random walk, clipped
        else:
            x = rng.choice(np.array([-32768, -513, -512, 0, 512, 513, 32767]), n)
        caps.append(np.clip(x, -32768, 32767).astype(np.int16))
        bauds.append(int(rng.choice(bauds_pool)))
        thrs.append(int(rng.choice([14000, 8000, 300, 0, 30000])))
    samples, offsets = A.modem._concat(caps)
    s = A.RxSession(offsets, bauds, thrs)
    s.upload(samples); s.run()
    _check_against_oracle(s.download(), caps, bauds, thrs)
    s.close()


def _device_round_trip(payloads, baud, tt=0.5):
    """encode on the GPU, decode the device-resident frames in place (ranges plan), return (RxBatch, TxSession)."""
    ts = int(baud * tt / 2)
    tx = A.TxSession(payloads, baud, ts)
    tx.upload(); tx.run()
    _cabi.stream_sync(0)
    rx = A.RxSession(tx.out_off[:-1].copy(), baud, 14000, lengths=tx.out_len.astype(np.int64))
    rx.bind(tx.d_out.ptr)
    rx.run()
    b = rx.download()
    return b, tx, rx


def test_config2_full_size_device_round_trip():
    """BASELINE config #2 at FULL size — 4096 captures x 1 KB at 1200 baud (2.47 G samples, 4.9 GB) —
    synthesized and decoded without leaving HBM: encode -> decode is the identity, every stage integer
    is the one SURVEY §8 derives (clock 0, 604 training + terminator bits, 14336 coded bits)."""
    rng = np.random.default_rng(22)
    B = 4096
    pls = [p.tobytes() for p in rng.integers(0, 256, size=(B, 1024), dtype=np.uint8)]
    b, tx, rx = _device_round_trip(pls, 1200)
    assert int(tx.out_len.sum()) == B * 602400
    assert (b.status == 0).all() and (b.clock == 0).all() and (b.nbits == 14336).all() and (b.nbytes == 1024).all()
    assert (b.train_end == 604 * 40).all()
    assert b.payloads() == pls
    rx.close(); tx.close()


def test_config4_full_size_device_round_trip():
    """BASELINE config #4 at full size: 64 captures x 64 KB at 300 baud (146.8 M samples each, 18.8 GB)."""
    rng = np.random.default_rng(44)
    B = 64
    pls = [p.tobytes() for p in rng.integers(0, 256, size=(B, 65536), dtype=np.uint8)]
    b, tx, rx = _device_round_trip(pls, 300)
    assert (tx.out_len == 146830080).all()
    assert (b.status == 0).all() and (b.clock == 0).all() and (b.nbits == 917504).all()
    assert b.payloads() == pls
    rx.close(); tx.close()


@pytest.mark.parametrize("baud", [1200, 2400, 6000, 300])
def test_hamming_corrects_one_flipped_bit_per_codeword(baud):
    """encode -> corrupt -> decode: ONE coded bit of EVERY 7-bit codeword is replaced by the opposite tone
    (a random position per codeword); ECC.decode (afskmodem.py:145-163) must return the payload unchanged.
    With a second flip in some codewords the (mis)correction must equal the oracle's."""
    import torch
    rng = np.random.default_rng([7, baud])
    B, P = 64, 256
    bf = 48000 // baud
    pls = [p.tobytes() for p in rng.integers(0, 256, size=(B, P), dtype=np.uint8)]
    ts = int(baud * 0.5 / 2)
    tx = A.TxSession(pls, baud, ts)
    tx.upload(); tx.run(); _cabi.stream_sync(0)
    host = tx.download().samples.copy()
    x = torch.from_numpy(host).cuda()
    ncw = 2 * P                                                       # codewords per capture
    first = 2 * ts + 4                                                # first coded bit (training + terminator)
    r = torch.from_numpy(rng.integers(0, 7, size=(B, ncw))).cuda()
    starts = torch.from_numpy(tx.out_off[:-1].copy()).cuda()[:, None] + (first + 7 * torch.arange(ncw, device="cuda")[None, :] + r) * bf
    idx = starts.reshape(-1, 1) + torch.arange(bf, device="cuda")[None, :]          # [B*ncw, bf]
    q = bf // 4
    pos = torch.arange(bf, device="cuda")
    mark = torch.where((pos // q) % 2 == 0, 32767, -32768).to(torch.int16)             # afskmodem.py:80-85
    space = torch.where(pos < bf // 2, 32767, -32768).to(torch.int16)                  # :68-77
    is_mark = x[idx[:, q]] < 0                                                         # second quarter low <=> mark
    x[idx] = torch.where(is_mark[:, None], space[None, :], mark[None, :])
    flipped = x.cpu().numpy()
    rx = A.RxSession(tx.out_off[:-1].copy(), baud, 14000, lengths=tx.out_len.astype(np.int64))
    rx.upload(flipped); rx.run()
    b = rx.download()
    assert (b.nbits == 14 * P).all()
    assert b.payloads() == pls, "single-bit errors must be corrected"
    # a second flip in the first 40 codewords of every capture: miscorrection, identical to the oracle's
    r2 = (r[:, :40] + 1 + torch.from_numpy(rng.integers(0, 6, size=(B, 40))).cuda()) % 7
    starts2 = torch.from_numpy(tx.out_off[:-1].copy()).cuda()[:, None] + (first + 7 * torch.arange(40, device="cuda")[None, :] + r2) * bf
    idx2 = starts2.reshape(-1, 1) + torch.arange(bf, device="cuda")[None, :]
    is_mark2 = x[idx2[:, q]] < 0
    x[idx2] = torch.where(is_mark2[:, None], space[None, :], mark[None, :])
    twice = x.cpu().numpy()
    rx.upload(twice); rx.run()
    b2 = rx.download()
    assert b2.payloads() != pls
    for i in (0, B // 2, B - 1):
        o = O.rx_decode(twice[int(tx.out_off[i]):int(tx.out_off[i]) + int(tx.out_len[i])], baud, 14000)
        assert (int(b2.nbits[i]), b2.payload(i)) == (o["nbits"], o["data"])
    rx.close(); tx.close()
