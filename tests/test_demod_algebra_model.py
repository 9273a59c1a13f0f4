"""CPU check of the integer closed forms the demodulator kernels rely on, against the
reference rule written out directly from afskmodem.py:287-296 (__amplify), :101-107 (getDiff) and
:342-351 (__decodeBit):

  * 2 * sum|T - amp| = 65535 * (bf - U) + T.(p + n) = 65535 * bf - 65534 * U + 2 * Xn
    with  U = T.(p - n),  Xn = T.n   (per template; sum T = 0)
  * the bit is 1 iff Um > Us, except on Um == Us where it is 1 iff Ns > Nm and floor(M/bf) < floor(S/bf)
  * the single-accumulator form of k_demod's merge mode / k_demod_lane: D = (mark - space) / 2 is
    0, -1, +1, 0 per quarter, D.c = (Um - Us) / 2 decides, and D.c == 0 with D.n >= 0 is a 0 without
    looking at the floors (D.n = (Nm - Ns) / 2)
  * the end detector: floor(sum|x| / bf) < thr  <=>  sum|x| < thr * bf

Windows are drawn to hit the limiter edges (|x| = 512 / 513), full scale, and pure-noise ties."""
import numpy as np
import pytest

HI, LO = 32767, -32768


def reference_bit(x: np.ndarray, bf: int) -> int:
    q = bf // 4
    amp = np.where(x > 512, HI, np.where(x < -512, LO, 0)).astype(np.int64)          # :287-296
    mark = np.array(([HI] * q + [LO] * q) * 2, dtype=np.int64)                         # getMarkTone :81-85
    space = np.array([HI] * (2 * q) + [LO] * (2 * q), dtype=np.int64)                  # getSpaceTone :68-78
    m = int(np.abs(mark - amp).sum() / bf)                                             # getDiff :107 (int(x / n))
    s = int(np.abs(space - amp).sum() / bf)
    return 1 if m < s else 0                                                           # :348-351


def closed_form_bit(x: np.ndarray, bf: int):
    q = bf // 4
    p = (x > 512).astype(np.int64)
    n = (x < -512).astype(np.int64)
    c = p - n
    quarter = np.arange(bf) // q
    Tm = np.where(quarter % 2 == 0, 1, -1)                   # mark  + - + -
    Ts = np.where(quarter < 2, 1, -1)                        # space + + - -
    Um, Us, Nm, Ns = int(Tm @ c), int(Ts @ c), int(Tm @ n), int(Ts @ n)
    # the closed form of the two sums themselves
    amp = np.where(x > 512, HI, np.where(x < -512, LO, 0)).astype(np.int64)
    for T, U, Xn in ((Tm, Um, Nm), (Ts, Us, Ns)):
        tone = np.where(T > 0, HI, LO)
        assert 2 * int(np.abs(tone - amp).sum()) == 65535 * bf - 65534 * U + 2 * Xn
        assert 65535 * bf - 65534 * U + 2 * Xn == 65535 * (bf - U) + int(T @ (p + n))
    M2 = 65535 * bf - 65534 * Um + 2 * Nm                    # 2 * sum|mark - amp|, as in the kernels
    S2 = 65535 * bf - 65534 * Us + 2 * Ns
    # plain-mode decision (k_demod <NT,false>, k_demod_shift)
    if Um != Us:
        plain = 1 if Um > Us else 0
    elif Ns > Nm:
        two_bf = 2 * bf
        plain = 1 if (S2 - M2 >= two_bf) or (M2 < (S2 // two_bf) * two_bf) else 0
    else:
        plain = 0
    # single-accumulator decision (merge mode, k_demod_lane): D weights 0,-1,+1,0 per quarter
    D = np.select([quarter == 1, quarter == 2], [-1, 1], 0)
    Dc, Dn = int(D @ c), int(D @ n)
    assert 2 * Dc == Um - Us and 2 * Dn == Nm - Ns
    if Dc != 0:
        single = 1 if Dc > 0 else 0
    elif Dn < 0:
        single = plain                                        # the kernels recompute the full sums here
    else:
        single = 0
    return plain, single


def windows(bf: int, rng, count: int):
    q = bf // 4
    mark = np.array(([HI] * q + [LO] * q) * 2)
    space = np.array([HI] * (2 * q) + [LO] * (2 * q))
    for k in range(count):
        kind = k % 6
        if kind == 0:
            x = rng.integers(-32768, 32768, bf)
        elif kind == 1:
            x = rng.choice([-513, -512, -511, 0, 511, 512, 513], bf)            # limiter edges
        elif kind == 2:
            x = (mark if rng.random() < 0.5 else space) * rng.choice([1.0, 0.3, 0.02]) + rng.normal(0, 9000, bf)
        elif kind == 3:
            x = rng.normal(0, rng.choice([300, 600, 5000]), bf)                 # noise only: ties are common
        elif kind == 4:
            x = rng.choice([LO, HI, 0], bf)
        else:
            x = np.where(rng.random(bf) < 0.5, mark, space) + rng.normal(0, 20000, bf)
        yield np.clip(np.round(x), -32768, 32767).astype(np.int64)


@pytest.mark.parametrize("bf", [4, 8, 12, 16, 20, 24, 32, 40, 48, 60, 80, 100, 120, 160, 200, 480, 2000])
def test_decision_closed_forms_equal_reference_rule(bf):
    rng = np.random.default_rng(bf)
    ties = 0
    for x in windows(bf, rng, 600 if bf <= 160 else 120):
        want = reference_bit(x, bf)
        plain, single = closed_form_bit(x, bf)
        assert plain == want and single == want, (bf, x.tolist())
        ties += 1
    assert ties


@pytest.mark.parametrize("bf", [8, 40, 160])
def test_end_detector_threshold_form(bf):
    rng = np.random.default_rng(1000 + bf)
    for x in windows(bf, rng, 300):
        a = int(np.abs(x).sum())
        for thr in (0, 1, 511, 8000, 14000, 32768, 65537):
            assert (int(a / bf) < thr) == (a < thr * bf)          # getAmplitude :94-98 vs the kernels' compare
