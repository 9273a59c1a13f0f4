"""CPU: the oracle (oracle/afsk_oracle.c) against the committed outputs of the real reference
(tests/golden, produced by tests/golden/make_golden.py from /root/reference/afskmodem.py)."""
import hashlib

import numpy as np
import pytest

from conftest import gate_multi_stream, load_gate_golden, load_gate_multi_golden, load_rx_golden, load_tx_golden
from oracle import oracle as O

RX = load_rx_golden()
TX = load_tx_golden()


@pytest.mark.parametrize("meta,x", RX, ids=[m["name"] for m, _ in RX])
def test_rx_decode_matches_reference(meta, x):
    o = O.rx_decode(x, meta["baud"], meta["amp_end"])
    if meta["ctor_exc"]:
        assert o["status"] == O.ST_EXC_BAUD
        assert list(O.EXC_TEXT[o["status"]]) == meta["ctor_exc"]
        return
    if meta["exc"]:
        assert o["status"] < 0 and list(O.EXC_TEXT[o["status"]]) == meta["exc"]
        return
    assert o["status"] >= 0
    assert o["data"].hex() == meta["data_hex"]
    assert o["clock"] == (-1 if meta["clock"] is None else meta["clock"])
    assert o["train_end"] == (-1 if meta["train_end"] is None else meta["train_end"])
    assert o["nbits"] == (meta["nbits"] or 0)
    assert o["nbytes"] == (meta["nbytes"] or 0)
    assert (o["status"] == O.ST_NO_CLOCK) == bool(meta["no_clock"])
    assert (o["status"] != O.ST_OK) == bool(meta["no_data"])


@pytest.mark.parametrize("meta,fr", TX, ids=[m["name"] for m, _ in TX])
def test_tx_frames_match_reference(meta, fr):
    mine = O.tx_frames(bytes.fromhex(meta["payload_hex"]), meta["baud"], meta["training_time"])
    if meta["exc"]:
        assert mine is None and meta["exc"] == ["Exception", "Invalid baud rate."]
        return
    assert len(mine) == meta["n"]
    assert hashlib.sha256(mine.astype("<i2").tobytes()).hexdigest() == meta["sha256"]
    if fr is not None:
        assert np.array_equal(mine, fr)


def test_tone_tables():
    # afskmodem.py:68-91 — lengths/values for the classes of baud in SURVEY F1/F2
    sp = O.tone("space", 1200)
    mk = O.tone("mark", 1200)
    assert list(sp) == [32767] * 20 + [-32768] * 20
    assert list(mk) == ([32767] * 10 + [-32768] * 10) * 2
    assert list(O.tone("training", 1200)) == list(mk) + list(sp)
    assert len(O.tone("mark", 4800)) == 8 and len(O.tone("space", 4800)) == 10
    assert O.tone("mark", 9600) is None and O.tone("space", 9600) is not None
    assert O.tone("space", 1100) is None


def test_ecc_roundtrip_and_single_error_correction():
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2, size=4 * 64, dtype=np.uint8)
    enc = O.ecc_encode(bits)
    assert len(enc) == 7 * 64
    assert np.array_equal(O.ecc_decode(enc), bits)
    for g in range(64):
        e = enc.copy()
        e[7 * g + int(rng.integers(0, 7))] ^= 1
        assert np.array_equal(O.ecc_decode(e), bits)
    assert len(O.ecc_decode(enc[:13])) == 4 and len(O.ecc_decode(enc[:6])) == 0


@pytest.mark.parametrize("g", load_gate_golden(), ids=[g["name"] for g in load_gate_golden()])
def test_listen_gate_matches_reference_receive(g):
    hello = dict((m["name"], x) for m, x in RX)["readme_hello_1200"]
    s = np.concatenate([np.zeros(g["lead"], np.int16), (hello.astype(np.float64) * g["gain"]).astype(np.int16),
                        np.zeros(g["tail_zeros"], np.int16)])
    ok, a, b = O.listen_gate(s, g["amp_start"], g["amp_end"], int(g["timeout"] * 48000))
    if g["exc"]:            # the stub stream ran dry before the gate opened: reference would block
        assert not ok
        return
    assert ok == (not g["timed_out"])
    if ok:
        assert b // 2048 == g["reads"]          # chunks consumed, including the discarded first
        o = O.rx_decode(s[a:b], 1200, g["amp_end"])
        assert o["data"].hex() == g["ret_hex"] and o["clock"] == g["clock"]
        assert o["train_end"] == g["train_end"] and o["nbits"] == g["nbits"]
    else:
        # reads = 1 discarded + chunks examined before the timeout fired
        assert g["ret_hex"] == ""


@pytest.mark.parametrize("g", load_gate_multi_golden(), ids=[g["name"] for g in load_gate_multi_golden()])
def test_listen_gate_multi_matches_successive_reference_receives(g):
    """Several transmissions in one recording: the oracle's walk equals successive receive() calls
    of ONE reference Receiver (which keeps its stream open between calls, afskmodem.py:283)."""
    s = gate_multi_stream(g, O.tx_frames)
    calls = O.listen_gate_multi(s, g["amp_start"], g["amp_end"], int(g["timeout"] * 48000))
    ref = g["calls"]
    assert len(calls) >= len(ref)
    for (rec, a, b), c in zip(calls, ref):
        assert rec == (not c["timed_out"])
        if not rec:
            assert c["ret_hex"] == ""
            continue
        assert b // 2048 == c["reads_after"] and a // 2048 > c["reads_before"]
        o = O.rx_decode(s[a:b], g["baud"], g["amp_end"])
        assert o["data"].hex() == c["ret_hex"]
        assert (o["clock"], o["train_end"], o["nbits"]) == (c["clock"], c["train_end"], c["nbits"])
    # finite-recording convention: the only call the reference does not report (its stream.read
    # would block) is a last recording still open when the samples run out
    extra = calls[len(ref):]
    assert len(extra) <= 1 and all(rec and b == (len(s) // 2048) * 2048 for rec, a, b in extra)
    # the single-call gate is the first call of the walk
    if calls:
        assert O.listen_gate(s, g["amp_start"], g["amp_end"], int(g["timeout"] * 48000)) == calls[0]
