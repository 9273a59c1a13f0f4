"""CPU: the N>1 path (capture sharding + host gather) on a world-size-2 gloo group.  The decode of
each shard is stood in for by the oracle (this is a test of the sharding/gather logic; on GPUs each
rank calls RxSession on its own device — bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from afskmodem_b200.shard import gather_batches, shard_captures
from oracle import oracle as O


def _make_caps():
    rng = np.random.default_rng(9)
    caps = []
    for i in range(11):
        pl = rng.integers(0, 256, int(rng.integers(1, 30)), dtype=np.uint8).tobytes()
        fr = O.tx_frames(pl, 1200, 0.05)
        x = np.clip(fr.astype(np.int32) + np.round(rng.normal(0, 5000, len(fr))).astype(np.int32),
                    -32768, 32767).astype(np.int16)
        caps.append(np.concatenate([np.zeros(int(rng.integers(0, 3000)), np.int16), x]))
    return caps


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    caps = _make_caps()
    lo, hi = shard_captures([len(c) for c in caps], world)[rank]
    local = [(i, O.rx_decode(caps[i], 1200, 14000)["data"]) for i in range(lo, hi)]
    allr = gather_batches(local)
    if rank == 0:
        flat = [x for part in allr for x in part]
        np.save(out_path, np.array([i for i, _ in flat]))
        with open(out_path + ".bin", "wb") as f:
            f.write(b"|".join(d.hex().encode() for _, d in flat))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "order.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    order = np.load(out)
    caps = _make_caps()
    assert list(order) == list(range(len(caps)))          # disjoint, ordered, complete
    got = open(out + ".bin", "rb").read().split(b"|")
    want = [O.rx_decode(c, 1200, 14000)["data"].hex().encode() for c in caps]
    assert got == want
