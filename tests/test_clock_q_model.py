"""CPU: the arithmetic of k_clock_q (afsk_rx.cu: clock recovery for bit lengths of 4 Q = 8..24 frames, 6000..2000 baud)
restated in numpy — Q-sample sums, their differences / sums, the distance D of the training cycle at every candidate, the
quotient by one multiply-high, the key (quotient << 12 | position), the block rotation and the masked ends — against the
reference rule (afskmodem.py:322-339: first index of the minimum of floor(sum |T - x| / 2bf) over 4096 - 2bf candidates)."""
import numpy as np
import pytest

from oracle import oracle as O

QS = [2, 3, 4, 5, 6]


@pytest.mark.parametrize("Q", QS)
def test_multiply_high_is_the_floor(Q):
    """floor(D / 2bf) == umulhi(D, ceil(2^32 / 2bf)) for EVERY distance the search can produce (D <= 65535 * 2bf < 2^22)"""
    div = 8 * Q
    magic = ((1 << 32) + div - 1) // div
    assert magic < 1 << 32
    dmax = 65535 * div
    assert dmax < 1 << 22
    D = np.arange(0, dmax + 1, dtype=np.uint64)
    assert np.array_equal((D * np.uint64(magic)) >> np.uint64(32), D // np.uint64(div))


def clock_q_model(y, e, Q):
    """y: the 16-byte aligned sample stream the kernel reads (the capture starts at y[e], e = 0..7); thread t of the CTA
    scores the 32 candidates at positions 32 b .. 32 b + 31, b = (t + 126) & 127, from 4 + Q vectors of its own"""
    span = 4096 - 8 * Q
    div = 8 * Q
    magic = ((1 << 32) + div - 1) // div
    c0 = 65535 * 4 * Q
    nA, nE = 32 + 7 * Q, 32 + 2 * Q
    best = 0xFFFFFFFF
    for t in range(128):
        blk = (t + 126) & 127
        masked = t < 32                                   # warp 0: blocks 126, 127, 0 .. 29
        W = np.zeros(8 * (4 + Q), np.int64)
        for r in range(4 + Q):
            v = 4 * blk + r
            if (not masked) or v < 512 or (v == 512 and e > 0):
                W[8 * r:8 * r + 8] = y[8 * v:8 * v + 8]
        a = np.array([W[j:j + Q].sum() for j in range(nA)])
        ee = a[:nE] - a[Q:Q + nE]
        bb = a[4 * Q:4 * Q + nE] + a[5 * Q:5 * Q + nE]
        k = np.arange(32)
        D = c0 - bb[k] + bb[k + 2 * Q] - ee[k] - ee[k + 2 * Q]
        key = ((D.astype(np.uint64) * np.uint64(magic)) >> np.uint64(32)).astype(np.int64) * 4096 + 32 * blk + k
        if masked:
            lo, hi = e - 32 * blk, e + span - 32 * blk
            key = np.where((k < lo) | (k >= hi), 0xFFFFFFFF, key)
        else:
            assert 0 <= D.min() and D.max() <= 65535 * div    # an unmasked thread never leaves the capture
            assert e <= 32 * blk and 32 * blk + 31 < e + span
        best = min(best, int(key.min()))
    return (best & 4095) - e


@pytest.mark.parametrize("Q", QS)
def test_key_minimum_is_the_reference_first_minimum(Q):
    bf = 4 * Q
    baud = 48000 // bf
    rng = np.random.default_rng([9, Q])
    for trial in range(16):
        fr = O.tx_frames(b"abc", baud, 0.12)
        lead = int(rng.choice([0, 1, bf - 1, 2 * bf + 7, 4096 - 2 * bf - 1, 4096 - 2 * bf, 4090])) if trial % 4 == 0 else int(rng.integers(0, 64))
        x = np.concatenate([np.zeros(lead, np.int16), fr]).astype(np.float64)
        if trial % 3 == 1:
            x = x + np.round(rng.normal(0, float(rng.choice([4000, 15000, 30000])), len(x)))
        if trial % 5 == 2:
            x = np.round(rng.normal(0, 12000, len(x)))
        if trial == 14:
            x = np.full(5000, -32768.0)
        if trial == 15:
            x = np.full(5000, 32767.0)
        x = np.clip(x, -32768, 32767).astype(np.int16)
        assert len(x) >= 4096 + 8
        want = O.rx_decode(x, baud, 14000)["clock"]
        for e in (0, 1, 3, 4, 7):
            # what surrounds the capture in the batch buffer (another capture's samples) must not matter
            y = np.concatenate([rng.integers(-32768, 32768, e), x[:4096].astype(np.int64), rng.integers(-32768, 32768, 64)])
            assert clock_q_model(y, e, Q) == want, (Q, trial, e)
