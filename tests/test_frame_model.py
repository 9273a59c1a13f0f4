"""CPU model of k_frame / k_frame_warp's packed-word arithmetic (afskmodem_b200/csrc/afsk_rx.cu) against the
reference loops restated directly:

  * phase 1 (afskmodem.py:362-366, __scanTraining :386-390): first k with bits[k-3..k] == 1,0,0,0, found on
    32-bit plane words as  M = (b << 3 | carry) & ~(b << 2 | ..) & ~(b << 1 | ..) & ~b  (funnel shifts with
    the previous word), the windows past K masked off;
  * phase 2 (:372-378): first quiet window at or after k0 = k + 1;
  * ECC.decode (:145-163) through the 128-entry nibble table and __bitsToBytes (:393-399): four bytes
    (56 coded bits) cut out of three plane words at an arbitrary bit offset.
"""
import numpy as np
import pytest

from oracle import oracle as O


def pack(bits: np.ndarray) -> np.ndarray:
    n = (len(bits) + 31) // 32 + 4                               # the kernels may touch up to 3 words past the end
    padded = np.zeros(n * 32, dtype=np.uint8)
    padded[:len(bits)] = bits
    return np.packbits(padded.reshape(-1, 32)[:, ::-1], axis=1).view(">u4").reshape(-1).astype(np.uint64)


def funnel_l(prev, cur, s):
    return ((cur << np.uint64(s)) | (prev >> np.uint64(32 - s))) & np.uint64(0xFFFFFFFF)


def funnel_r(lo, hi, s):
    return (((hi << np.uint64(32)) | lo) >> np.uint64(s)) & np.uint64(0xFFFFFFFF)


def hamming_nibble(cw: int) -> int:
    c = [(cw >> i) & 1 for i in range(7)]
    s0, s1, s2 = c[0] ^ c[2] ^ c[4] ^ c[6], c[1] ^ c[2] ^ c[5] ^ c[6], c[3] ^ c[4] ^ c[5] ^ c[6]
    e = 4 * s2 + 2 * s1 + s0
    if e:
        c[e - 1] ^= 1
    return (c[2] << 3) | (c[4] << 2) | (c[5] << 1) | c[6]


LUT = [hamming_nibble(i) for i in range(128)]


def kernel_model(bits: np.ndarray, quiet: np.ndarray):
    K = len(bits)
    bw, qw = pack(bits), pack(quiet)
    nwords = (K + 31) // 32
    NONE = 0x7FFFFFFF
    kterm = NONE
    for j in range(nwords):
        cur, prev = bw[j], (bw[j - 1] if j else np.uint64(0))
        M = funnel_l(prev, cur, 3) & ~funnel_l(prev, cur, 2) & ~funnel_l(prev, cur, 1) & ~cur & np.uint64(0xFFFFFFFF)
        rem = K - 32 * j
        if rem < 32:
            M &= np.uint64((1 << rem) - 1)
        if M:
            kterm = 32 * j + (int(M) & -int(M)).bit_length() - 1
            break
    k0 = K if kterm == NONE else kterm + 1
    k1 = K
    for j in range(k0 >> 5, nwords):
        M = qw[j]
        if j == (k0 >> 5):
            M &= ~np.uint64((1 << (k0 & 31)) - 1) & np.uint64(0xFFFFFFFF)
        rem = K - 32 * j
        if rem < 32:
            M &= np.uint64((1 << rem) - 1)
        if M:
            k1 = 32 * j + (int(M) & -int(M)).bit_length() - 1
            break
    nbits = k1 - k0
    nbytes = (nbits // 7) // 2
    out = bytearray()
    for i in range(nbytes >> 2):                                  # four bytes per step, three plane words
        pos = k0 + 56 * i
        wi, sh = pos >> 5, pos & 31
        w0, w1, w2 = bw[wi], bw[wi + 1], bw[wi + 2]
        for jb in range(4):
            sft = sh + 14 * jb
            a = w0 if sft < 32 else (w1 if sft < 64 else w2)
            b = w1 if sft < 32 else (w2 if sft < 64 else np.uint64(0))
            v14 = int(funnel_r(a, b, sft & 31)) & 0x3FFF
            out.append((LUT[v14 & 0x7F] << 4) | LUT[(v14 >> 7) & 0x7F])
    for i in range(4 * (nbytes >> 2), nbytes):
        pos = k0 + 14 * i
        v14 = int(funnel_r(bw[pos >> 5], bw[(pos >> 5) + 1], pos & 31)) & 0x3FFF
        out.append((LUT[v14 & 0x7F] << 4) | LUT[(v14 >> 7) & 0x7F])
    return k0, nbits, bytes(out)


def reference_model(bits: np.ndarray, quiet: np.ndarray):
    K = len(bits)
    k0, reg = K, [0, 0, 0, 0]
    for k in range(K):                                            # :362-366 with the shift register of :386-390
        reg = reg[1:] + [int(bits[k])]
        if reg == [1, 0, 0, 0]:
            k0 = k + 1
            break
    k1 = K
    for k in range(k0, K):                                        # :372-378
        if quiet[k]:
            k1 = k
            break
    coded = np.asarray(bits[k0:k1], dtype=np.uint8)
    data = O.ecc_decode(coded)                                    # ECC.decode :154-163 (C restatement, pinned by the goldens)
    nbytes = len(data) // 8
    by = np.packbits(data[:nbytes * 8]).tobytes() if nbytes else b""
    return k0, len(coded), by


@pytest.mark.parametrize("seed", range(12))
def test_packed_framing_equals_reference_loops(seed):
    rng = np.random.default_rng(seed)
    for trial in range(25):
        K = int(rng.integers(0, 700))
        lead = int(rng.integers(0, 90))
        body = int(rng.integers(0, 500))
        bits = np.concatenate([np.tile([1, 0], lead // 2 + 1)[:lead], [1, 0, 0, 0] if trial % 5 else [],
                               rng.integers(0, 2, body)]).astype(np.uint8)[:K]
        bits = np.concatenate([bits, rng.integers(0, 2, max(0, K - len(bits))).astype(np.uint8)])
        quiet = (rng.random(K) < (0.0 if trial % 3 == 0 else 0.01)).astype(np.uint8)
        assert kernel_model(bits, quiet) == reference_model(bits, quiet), (seed, trial, K)


def test_nibble_table_equals_oracle_ecc():
    for cw in range(128):
        c = np.array([(cw >> i) & 1 for i in range(7)], dtype=np.uint8)
        want = O.ecc_decode(c)
        got = LUT[cw]
        assert [(got >> 3) & 1, (got >> 2) & 1, (got >> 1) & 1, got & 1] == list(want), cw
