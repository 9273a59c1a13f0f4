import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import afskmodem_b200 as A
from afskmodem_b200 import _cabi
A.LOG_LEVEL = 5
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
corpus = bench.Corpus(wl, bench.resolve_workload(wl)[2])
B = corpus.total
bench.Ctx.local = 0; bench.Ctx.dev = torch.device("cuda", 0)
samples, offsets = bench.build_on_gpu(corpus, 0, B, 0)
sess = A.RxSession(offsets, corpus.baud_rx, corpus.amp_end, 0)
sess.bind(samples.data_ptr(), samples.numel())
L = _cabi.lib()
for fused in (1, 0):
    L.afsk_rx_plan_set_option(sess.plan, _cabi.OPT_FUSED, fused)
    for _ in range(3):
        sess.run()
    torch.cuda.synchronize()
    buf = np.zeros((4096, 10), dtype=np.uint64)
    L.afsk_dbg_trace_read(buf.ctypes.data_as(C.c_void_p), 4096, 1)
    sess.run(); torch.cuda.synchronize()
    n = L.afsk_dbg_trace_read(buf.ctypes.data_as(C.c_void_p), 4096, 1)
    r = buf[:n].astype(np.int64)
    t0 = r[:, 2][r[:, 2] > 0].min() if n else 0
    fin = r[(r[:, 0] & 255) == 6]
    if len(fin):
        cta = fin[:, 0] >> 8
        tt = (fin[:, 2] - t0) / 1e3
        order = np.argsort(cta)
        print("aux exit per CTA (us):", " ".join(f"{int(cta[i])}:{tt[i]:.0f}" for i in order[:296:4]))
        fr = r[(r[:, 0] & 255) == 2]
        cnt = np.bincount((fr[:, 0] >> 8).astype(int), minlength=296)
        print("sampled frame jobs per CTA:", cnt.tolist())
    for kind, name in ((1, "clock job"), (2, "frame job (fused)"), (3, "frame (k_frame_warp)"), (5, "consumer warp 0 finish"), (6, "aux exit")):
        k = r[(r[:, 0] & 255) == kind]
        if not len(k):
            continue
        nm = int(k[0, 1])
        d = np.diff(k[:, 2:2 + nm], axis=1) / 1e3
        print(f"{wl} fused={fused} {name}: {len(k)} records; stage durations us median {np.median(d, axis=0).round(1)} p90 {np.percentile(d, 90, axis=0).round(1)}; "
              f"start offsets us min {((k[:, 2] - t0) / 1e3).min():.0f} median {np.median((k[:, 2] - t0) / 1e3):.0f} max {((k[:, 2] - t0) / 1e3).max():.0f}", flush=True)
