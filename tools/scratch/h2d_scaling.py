"""Concurrent pinned H2D from every rank: default pinned vs write-combined pinned memory."""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
from afskmodem_b200 import _cabi
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
stride = int(os.environ.get("DEV_STRIDE", "1")); local = local * stride
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 1 << 30   # int16 elements = 2 GiB
d = _cabi.DeviceBuffer(local, 2 * N)
for mode in ("default", "default"):
    pin = _cabi.PinnedArray((N,), np.int16)   # a write-combined variant (cudaHostAllocWriteCombined) measured identical
    t0 = time.perf_counter(); pin.array[:] = 1; fill = time.perf_counter() - t0
    d.upload(pin.array); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): d.upload(pin.array)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"stride {stride} {mode}: world {world} max-rank H2D {ms.item():.1f} ms -> {2*N/ms.item()/1e6:.1f} GB/s per GPU, {world*2*N/ms.item()/1e6:.1f} GB/s total; cpu fill {fill*1e3:.0f} ms", flush=True)
    pin.close()
    if world > 1: dist.barrier()
if world > 1: dist.destroy_process_group()
