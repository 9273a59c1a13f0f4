"""CPU, build container only: live randomized comparison of the oracle with the UNMODIFIED
reference.  Skipped wherever /root/reference is absent (e.g. the GPU box)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_harness as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference not mounted")

BAUDS = [300, 600, 800, 1200, 2400, 4000, 6000, 12000, 4800, 960, 100]


@pytest.mark.parametrize("seed", range(4))
def test_random_trials(seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        baud = int(rng.choice(BAUDS))
        pl = rng.integers(0, 256, size=int(rng.integers(0, 16)), dtype=np.uint8).tobytes()
        tt = float(rng.choice([0.2, 0.1, 0.02]))
        fr = R.ref_save(pl, baud, tt)
        assert np.array_equal(fr, O.tx_frames(pl, baud, tt))
        lead = int(rng.integers(0, 5000)) if rng.random() < 0.5 else 0
        gain = float(rng.choice([1.0, 0.7, 0.45, 0.3]))
        sigma = float(rng.choice([0, 2000, 8000, 14000, 19000, 26000]))
        x = np.concatenate([np.zeros(lead, np.int16), fr]).astype(np.float64) * gain
        if sigma > 0:
            x = x + np.round(rng.normal(0, sigma, size=len(x)))
        x = np.clip(x, -32768, 32767).astype(np.int16)
        if rng.random() < 0.2:
            x = x[:int(rng.integers(1000, len(x)))]
        a0, a1 = [(18000, 14000), (14000, 11000), (9000, 8000)][int(rng.integers(0, 3))]
        r = R.ref_load(x, baud, a0, a1)
        o = O.rx_decode(x, baud, a1)
        if r["exc"]:
            assert o["status"] < 0 and O.EXC_TEXT[o["status"]] == r["exc"]
            continue
        assert r["ret"] == o["data"]
        assert (-1 if r["clock"] is None else r["clock"]) == o["clock"]
        assert (-1 if r["train_end"] is None else r["train_end"]) == o["train_end"]
        assert (r["nbits"] or 0) == o["nbits"] and (r["nbytes"] or 0) == o["nbytes"]
