import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import afskmodem_b200 as A
from afskmodem_b200 import _cabi
from test_gpu_round2 import _mixed_corpus, _decode_with
A.LOG_LEVEL = 5
for bauds in ((1200,), (6000,), (300,), (2400,), (300, 600, 1200, 2400, 4000, 6000, 1500, 800)):
    caps, baud, thr, pls = _mixed_corpus(51, 60, bauds=bauds)
    samples, offsets = A.modem._concat(caps)
    three, _ = _decode_with(samples, offsets, baud, thr, 0)
    for mode, fk in (("fused clock only", 2), ("fused", 0)):
        for rep in range(3):
            f, _ = _decode_with(samples, offsets, baud, thr, 1, frame_kernel=fk)
            same_res = np.array_equal(three.results, f.results)
            bad = [i for i in range(len(caps)) if three.payload(i) != f.payload(i)]
            print(bauds, mode, rep, "results equal:", same_res, "payload mismatches:", len(bad), bad[:10], flush=True)
            if bad:
                i = bad[0]
                a, b = three.payload(i), f.payload(i)
                d = [k for k in range(len(a)) if a[k] != b[k]]
                print("   capture", i, "baud", baud[i], "nbytes", len(a), "first diffs at", d[:12], "of", len(d), a[:8].hex(), b[:8].hex())
