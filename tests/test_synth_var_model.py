"""CPU: the lane step of k_synth_var for tones of 8 frames and more (afsk_tx.cu), restated in Python — locate the lane's
first frame, list the nine bits its 64 frames can touch (inside the coded bits: symbols cut from one 21-bit stream of the
three codewords involved, read from two payload bytes), emit the 32 frame pairs with "the next bit has started" as a select
and the tone phase as a bit of the tone's mask — against the reference's frames (afskmodem.py:452-469, :239-244) for the
unequal-tone rates (mark tone two frames shorter than the space tone)."""
import numpy as np
import pytest

from oracle import oracle as O

HI, LO = 32767, -32768


def hamming(v):
    d0, d1, d2, d3 = (v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1
    return (d0 ^ d1 ^ d3) | ((d0 ^ d2 ^ d3) << 1) | (d0 << 2) | ((d1 ^ d2 ^ d3) << 3) | (d1 << 4) | (d2 << 5) | (d3 << 6)

def below(n): return 0xFFFFFFFF if n >= 32 else (1 << n) - 1

def synth(pay, baud, ts_cycles):
    bf = 48000 // baud
    q, h = bf >> 2, bf >> 1
    ml, sl = 4 * q, bf
    assert ml != sl and min(ml, sl) >= 8
    ts_bits = 2 * ts_cycles
    G = 2 * len(pay)
    cws = [hamming((pay[g >> 1] >> 4) if g % 2 == 0 else (pay[g >> 1] & 15)) for g in range(G)]
    ns = np.zeros(G + 1, np.int64)
    for g in range(G):
        ones = bin(cws[g]).count("1")
        ns[g + 1] = ns[g] + ones * ml + (7 - ones) * sl
    total_bits = ts_bits + 4 + 7 * G
    T0 = ts_cycles * (ml + sl)
    C0 = T0 + ml + 3 * sl
    E = C0 + ns[G]
    total_frames = E + 4800
    out_len = total_frames & ~1
    nvec = (out_len + 7) >> 3
    out = np.zeros(nvec * 8, np.int32)
    use_masks = bf <= 64
    mask_space = below((h + 1) >> 1)
    mask_mark = below((q + 1) >> 1) | (below((3 * q + 1) >> 1) & ~below(q) & 0xFFFFFFFF)

    def tx_bit(b):
        if b < ts_bits: return (~b) & 1
        t = b - ts_bits
        if t < 4: return 1 if t == 0 else 0
        j = t - 4
        return (cws[j // 7] >> (j % 7)) & 1

    def locate(n):   # -> (b, u, g, r): as var_locate
        if n < T0:
            cyc = n // (ml + sl); r = n - cyc * (ml + sl)
            return (2 * cyc + (1 if r >= ml else 0), r - ml if r >= ml else r, 0, 0)
        t = n - T0
        if t < ml: return (ts_bits, t, 0, 0)
        t -= ml
        if t < 3 * sl: k = t // sl; return (ts_bits + 1 + k, t - k * sl, 0, 0)
        t -= 3 * sl
        if G == 0 or t >= ns[G]: return (total_bits, 0, 0, 0)
        lo = int(np.searchsorted(ns, t, side='right')) - 1
        rem = t - ns[lo]; cw = cws[lo]; r = 0
        while r < 6:
            L = ml if (cw >> r) & 1 else sl
            if rem < L: break
            rem -= L; r += 1
        return (ts_bits + 4 + 7 * lo + r, int(rem), lo, r)

    first_coded = ts_bits + 4
    for vl in range(0, nvec, 8):          # one lane step: 8 vectors = 64 frames from F
        F = 8 * vl
        b0, u0, g0, r0 = locate(F)
        # ---- the lane's bit list: 9 bits from b0 (symbol, start relative to F)
        syms = [0] * 10
        if b0 >= first_coded and b0 + 8 < total_bits:
            # inside the coded bits: the (at most three) codewords as one 21-bit stream
            i0 = g0 >> 1
            B0 = pay[i0]; B1 = pay[i0 + 1] if i0 + 1 < len(pay) else 0
            B01 = (B0 << 8) | B1
            stream = 0
            for k in range(3):
                idx = (g0 & 1) + k
                nib = (B01 >> (12 - 4 * idx)) & 15
                stream |= hamming(nib) << (7 * k)
            for i in range(9): syms[i] = (stream >> (r0 + i)) & 1
        else:
            for i in range(9): syms[i] = tx_bit(b0 + i) if b0 + i < total_bits else 2
        starts = [0] * 10
        r = -u0
        for i in range(9):
            starts[i] = r
            r += (1 << 20) if syms[i] == 2 else (ml if syms[i] else sl)
        starts[9] = 1 << 21; syms[9] = 2                       # sentinel
        # ---- 32 frame pairs, branch-free: advance while the next bit has started
        cur = 0
        for k in range(8):
            for j in range(4):
                o = 8 * k + 2 * j
                adv = 1 if o >= starts[cur + 1] else 0        # tones of >= 8 frames: at most one advance per pair... per vector
                cur += adv
                assert not (o >= starts[cur + 1]), "two starts inside one pair step"
                sym = syms[cur]; ph = o - starts[cur]
                if sym == 2: val = 0
                elif use_masks: val = HI if ((mask_mark if sym else mask_space) >> ((ph >> 1) & 31)) & 1 else LO
                else:
                    hi = (ph < q or (2 * q <= ph < 3 * q)) if sym else (ph < h)
                    val = HI if hi else LO
                if F + o + 1 < len(out) + 1 and F + o < len(out):
                    out[F + o] = val; out[F + o + 1] = val
    return out[:out_len].astype(np.int16)


CASES = [(4800, 300, 0.5), (4800, 0, 0.5), (4800, 1, 0.0), (4800, 257, 0.013), (1600, 200, 0.1), (960, 100, 0.02),
         (320, 40, 0.1), (192, 20, 0.05), (4800, 50, 0.0021), (1600, 3, 0.0)]


@pytest.mark.parametrize("baud,n,tt", CASES)
def test_lane_walk_equals_reference_frames(baud, n, tt):
    rng = np.random.default_rng([31, baud, n])
    pay = bytes(rng.integers(0, 256, n, dtype=np.uint8))
    want = O.tx_frames(pay, baud, tt)
    got = synth(pay, baud, O.ts_cycles(baud, tt))
    assert len(got) == len(want)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("baud", [4800, 960])
def test_lane_walk_shortest_and_longest_codewords(baud):
    for pay in (b"\xff" * 64, b"\x00" * 64, b"\xff\x00" * 40):
        assert np.array_equal(synth(pay, baud, O.ts_cycles(baud, 0.05)), O.tx_frames(pay, baud, 0.05))
