"""CPU: the arithmetic of k_clock_q (afsk_rx.cu) restated in numpy — window sums a[j] of Q samples on the 16-byte
aligned stream, e = a[j] - a[j+Q], b = a[j] + a[j+Q], D_j = c0 - e[j] - e[j+2Q] - b[j+4Q] + b[j+6Q], the exact
multiply-high floor by 2bf = 8Q, the ((floor << 12) | position) key with positions outside [e, e + span) masked —
against the reference rule (afskmodem.py:322-339: first index of the minimum of int(sum|T - x| / 2bf))."""
import numpy as np
import pytest

from oracle import oracle as O

QS = [2, 3, 4, 5, 6]


@pytest.mark.parametrize("Q", QS)
def test_multiply_high_is_the_floor(Q):
    d = 8 * Q
    magic = ((1 << 32) + d - 1) // d
    assert magic < 1 << 32
    dmax = 65535 * d                       # sum |T - x| over 2bf samples
    assert dmax < 1 << 22
    D = np.arange(0, dmax + 1, dtype=np.uint64)
    assert np.array_equal((D * np.uint64(magic)) >> np.uint64(32), D // np.uint64(d))


def clock_q_model(x, e, Q):
    """x: samples of the buffer from the aligned address (the capture starts at x[e]); returns the clock index"""
    bf, span = 4 * Q, 4096 - 8 * Q
    d = 8 * Q
    magic = ((1 << 32) + d - 1) // d
    y = np.zeros(4096 + 8 * Q + 64, np.int64)
    nvalid = 4096 if e == 0 else 4104      # vectors 0..511, and vector 512 when the start is not aligned
    y[:nvalid] = x[:nvalid]
    P = np.concatenate([[0], np.cumsum(y)])
    j = np.arange(4096)
    a = lambda k: P[k + Q] - P[k]          # noqa: E731
    ee = lambda k: a(k) - a(k + Q)         # noqa: E731
    bb = lambda k: a(k + 4 * Q) + a(k + 5 * Q)   # noqa: E731
    D = 65535 * bf - bb(j) + bb(j + 2 * Q) - ee(j) - ee(j + 2 * Q)
    valid = (j >= e) & (j < e + span)
    assert D[valid].min() >= 0 and D[valid].max() < 1 << 22
    D32 = D.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    key = ((((D32 * np.uint64(magic)) >> np.uint64(32)) * np.uint64(4096) + j.astype(np.uint64)) & np.uint64(0xFFFFFFFF))
    key[~valid] = 0xFFFFFFFF
    return int(key.min() & np.uint64(4095)) - e


@pytest.mark.parametrize("Q", QS)
def test_key_minimum_is_the_reference_first_minimum(Q):
    bf = 4 * Q
    baud = 48000 // bf
    rng = np.random.default_rng([17, Q])
    for trial in range(24):
        e = trial % 8
        fr = O.tx_frames(b"abc", baud, 0.12)
        x = np.concatenate([np.zeros(int(rng.integers(0, 2 * bf + 50)), np.int16), fr]).astype(np.float64)
        if trial % 3 == 1:
            x = x + np.round(rng.normal(0, 15000, len(x)))
        if trial % 3 == 2:
            x = np.round(rng.normal(0, 12000, len(x)))
        if trial == 23:
            x = np.full(5000, -32768.0)
        if trial == 22:
            x = np.full(5000, 32767.0)
        x = np.clip(x, -32768, 32767).astype(np.int16)
        if len(x) < 4200:
            x = np.concatenate([x, np.zeros(4200 - len(x), np.int16)])
        # what precedes the capture in the buffer is somebody else's samples
        buf = np.concatenate([rng.integers(-32768, 32768, e).astype(np.int16), x])
        assert clock_q_model(buf, e, Q) == O.rx_decode(x, baud, 14000)["clock"], (Q, trial)
