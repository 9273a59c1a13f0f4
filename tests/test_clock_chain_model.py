"""CPU model of k_clock's candidate enumeration (afskmodem_b200/csrc/afsk_rx.cu): the candidates
i in [0, 4096 - 2 bf) are walked in chains i, i + q, i + 2q, ... (q = bf / 4) of at most kClockChain
candidates, tiled as items (block, residue), thread t taking items t, t + 128, ...  The model restates
that index arithmetic exactly (chain starts, valid-prefix length kv, first qualifying k) and must give
the reference's answer — first index of the minimum of floor(D / 2 bf), afskmodem.py:322-339 — on every
decodable bit length.  It guards the arithmetic the kernel depends on; the kernel itself is compared with
the oracle in test_gpu_parity.py."""
import numpy as np
import pytest

CHAIN, THREADS, SYNC = 35, 128, 4096          # kClockChain, kClockThreads, AFSK_SYNC_FRAMES

BIT_FRAMES = [bf for bf in range(4, 2048, 4) if 48000 % bf == 0]     # 48000 / baud with equal tone lengths


def chain_clock(x: np.ndarray, bf: int) -> int:
    q, span = bf // 4, SYNC - 2 * bf
    P = np.concatenate([[0], np.cumsum(x[:SYNC + 8].astype(np.int64))])
    c0, div = 65535 * bf, 2 * bf

    def D(i):
        return c0 + P[i] + P[i + 8 * q] - 2 * (P[i + q] - P[i + 2 * q] + P[i + 3 * q] - P[i + 4 * q] + P[i + 6 * q])

    qL = q * CHAIN
    nitems = ((span + qL - 1) // qL) * q
    seen = np.zeros(span, dtype=np.int32)
    best, per_thread = None, []
    for tid in range(THREADS):
        s0 = (tid % q) + qL * (tid // q)
        kv = int((span - s0 + q - 1) / q) if tid < nitems else 0          # C division truncates toward zero
        kv = min(max(kv, 0), CHAIN)
        Dv = []
        for k in range(CHAIN):
            if k < kv:
                d = int(D(s0 + k * q))
                seen[s0 + k * q] += 1
                best = d if best is None else min(best, d)
            else:
                d = -1                                                      # undefined in the kernel: any value
            Dv.append(d)
        per_thread.append((s0, kv, Dv))
    extra = []
    for it in range(THREADS, nitems):
        s1 = (it % q) + qL * (it // q)
        for k in range(CHAIN):
            i = s1 + k * q
            if i < span:
                seen[i] += 1
                extra.append(i)
                best = min(best, int(D(i)))
    assert np.all(seen == 1), "every candidate is scored exactly once"
    T = (best // div + 1) * div - 1
    first = None
    for s0, kv, Dv in per_thread:
        kf = CHAIN
        for k in range(CHAIN - 1, -1, -1):
            if Dv[k] <= T:
                kf = k
        if kf < kv:
            first = s0 + kf * q if first is None else min(first, s0 + kf * q)
    for i in extra:
        if int(D(i)) <= T:
            first = i if first is None else min(first, i)
    return first


def reference_clock(x: np.ndarray, bf: int) -> int:
    """__recoverClockIndex restated directly: first index of the minimum of floor(sum |T - x| / 2 bf)."""
    q = bf // 4
    hi, lo = 32767, -32768
    T = np.array([hi] * q + [lo] * q + [hi] * q + [lo] * q + [hi] * (2 * q) + [lo] * (2 * q), dtype=np.int64)
    span = SYNC - 2 * bf
    xs = x[:SYNC].astype(np.int64)
    win = np.lib.stride_tricks.sliding_window_view(xs, 2 * bf)[:span]
    d = np.abs(win - T).sum(axis=1) // (2 * bf)
    return int(np.argmin(d))


@pytest.mark.parametrize("bf", BIT_FRAMES)
def test_chain_enumeration_equals_reference(bf):
    rng = np.random.default_rng(bf)
    q = bf // 4
    cyc = np.array([32767] * q + [-32768] * q + [32767] * q + [-32768] * q + [32767] * (2 * q) + [-32768] * (2 * q))
    for trial in range(3):
        x = rng.integers(-32768, 32768, 5000).astype(np.int16)
        if trial:                                    # a training sequence at a random offset, noisy or clean
            o = int(rng.integers(0, SYNC))
            seg = np.tile(cyc, 5000 // len(cyc) + 1)[:5000 - o].astype(np.float64)
            seg = seg * float(rng.choice([1.0, 0.4])) + (rng.normal(0, 9000, len(seg)) if trial == 2 else 0)
            x[o:] = np.clip(seg, -32768, 32767).astype(np.int16)
        assert chain_clock(x, bf) == reference_clock(x, bf), (bf, trial)
