"""CPU: the arithmetic of the fused kernel's clock jobs (aux_clock_index in afsk_rx.cu) restated in numpy —
the multiply-shift floor, the (floor << 12 | index) key, the two passes over the 4096-frame window — against
the reference rule (afskmodem.py:322-339: first index of the minimum of floor(sum|T - x| / 2bf))."""
import numpy as np
import pytest

from oracle import oracle as O

BFS = [4, 8, 12, 16, 20, 24, 32, 40, 48, 60, 64, 80, 96, 100, 120, 128, 160]


def magic_shift(bf):
    d = 2 * bf
    l = 0
    while (1 << l) < d:
        l += 1
    shift = 28 + l
    return ((1 << shift) + d - 1) // d, shift


@pytest.mark.parametrize("bf", BFS + [200, 480, 1000, 2000])
def test_multiply_shift_is_the_floor(bf):
    """exact for every D the clock search can produce (D <= 65535 * 2bf < 2^28), checked on the boundaries
    of every quotient and on random values"""
    m, s = magic_shift(bf)
    assert m < 1 << 32
    d = 2 * bf
    dmax = 65535 * d
    assert dmax < 1 << 28
    q = np.arange(0, 65536, dtype=np.uint64)
    for D in (q * d, q * d + (d - 1), np.minimum(q * d + 1, dmax)):
        assert np.array_equal((D * np.uint64(m)) >> np.uint64(s), D // np.uint64(d))
    rng = np.random.default_rng(bf)
    D = rng.integers(0, dmax + 1, 2_000_000).astype(np.uint64)
    assert np.array_equal((D * np.uint64(m)) >> np.uint64(s), D // np.uint64(d))


def aux_clock_model(x, bf):
    """numpy restatement of aux_clock_index: passes of 2048 candidates, D from prefix taps, key minimum"""
    q, span = bf // 4, 4096 - 2 * bf
    m, s = magic_shift(bf)
    c0 = 65535 * bf
    best = 0xFFFFFFFF
    for s0 in range(0, span, 2048):
        C = min(2048, span - s0)
        y = x[s0:s0 + C + 2 * bf].astype(np.int64)
        P = np.concatenate([[0], np.cumsum(y)])
        i = np.arange(C)
        t = lambda k: P[i + k * q]      # noqa: E731
        D = c0 + t(0) + t(8) - 2 * (t(1) - t(2) + t(3) - t(4) + t(6))
        assert D.min() >= 0 and D.max() < 1 << 28
        fl = (D.astype(np.uint64) * np.uint64(m)) >> np.uint64(s)
        key = (fl.astype(np.int64) << 12) | (s0 + i)
        best = min(best, int(key.min()))
    return best & 4095


@pytest.mark.parametrize("bf", BFS)
def test_key_minimum_is_the_reference_first_minimum(bf):
    baud = 48000 // bf
    rng = np.random.default_rng([3, bf])
    for trial in range(12):
        fr = O.tx_frames(b"abc", baud, 0.3)
        x = np.concatenate([np.zeros(int(rng.integers(0, 2 * bf + 50)), np.int16), fr]).astype(np.float64)
        if trial % 3 == 1:
            x = x + np.round(rng.normal(0, 15000, len(x)))
        if trial % 3 == 2:
            x = np.round(rng.normal(0, 12000, len(x)))
        if trial == 11:
            x = np.full(5000, -32768.0)
        x = np.clip(x, -32768, 32767).astype(np.int16)
        if len(x) < 4096:
            x = np.concatenate([x, np.zeros(4096 - len(x), np.int16)])
        assert aux_clock_model(x, bf) == O.rx_decode(x, baud, 14000)["clock"], (bf, trial)
