"""Multi-GPU sharding: captures are independent (Receiver.load touches one file,
afskmodem.py:420-430), so a corpus is cut into contiguous capture ranges — one range per rank/GPU —
with no collective on the data path.  Results (a few bytes per capture) are gathered on the host at
the end, in corpus order, into ONE RxBatch.

Ranges are balanced by PREDICTED TIME, not raw samples: the demodulator's throughput depends on the
bit length (DEMOD_GBS, measured per baud on a B200, profiles/), and every capture carries a fixed cost
for clock recovery and framing.  For a host-buffer decode the PCIe copy dominates and is uniform per
sample, which `capture_cost(..., resident=False)` models.
"""
from __future__ import annotations

import numpy as np

# k_demod* throughput in GB/s of int16 samples by samples-per-bit (48000 / baud), B200, device-resident
# (tools/baud_sweep.py; profiles/r2_baud_sweep.json).  Bit lengths that are not in the table: DEMOD_GBS_DEFAULT.
DEMOD_GBS = {200: 6960, 160: 6399, 128: 6235, 120: 6924, 100: 7032, 96: 6826, 80: 6370, 64: 6285, 60: 6953, 48: 7092,
             40: 7029, 32: 5859, 24: 6812, 20: 6384, 16: 6774, 12: 6283, 8: 6551, 4: 2902}
DEMOD_GBS_DEFAULT = 5000.0
PER_CAPTURE_NS = 6.0           # clock recovery + framing per capture (c2: 35 us / 4096, c3: 68 us / 16384)
PCIE_GBS = 55.0                # pinned host -> device, one B200 on PCIe 5 x16


def capture_cost(lengths, baud=None, resident: bool = True) -> np.ndarray:
    """Predicted nanoseconds per capture.  ``baud``: scalar or array (captures whose baud the reference
    rejects cost only the fixed part: they are never demodulated).  resident=False: the samples come
    from host memory, so every sample also crosses the PCIe link."""
    lengths = np.asarray(lengths, dtype=np.float64)
    gbs = np.full(len(lengths), DEMOD_GBS_DEFAULT)
    if baud is not None:
        b = np.broadcast_to(np.asarray(baud, dtype=np.int64), lengths.shape)
        ok = (b > 0) & (48000 % np.where(b > 0, b, 1) == 0)
        bf = np.where(ok, 48000 // np.where(b > 0, b, 1), 0)
        gbs = np.array([DEMOD_GBS.get(int(v), DEMOD_GBS_DEFAULT) for v in bf], dtype=np.float64)
        decodable = ok & (bf % 4 == 0)                         # others raise, nothing is streamed
        gbs = np.where(decodable, gbs, np.inf)
    per_sample = 2.0 / gbs                                      # ns per sample (GB/s == bytes/ns)
    if not resident:
        per_sample = np.maximum(per_sample, 2.0 / PCIE_GBS)     # the copy and the kernels overlap
    return lengths * per_sample + PER_CAPTURE_NS


def shard_captures(lengths, world_size: int, weights=None) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) capture ranges, one per rank, balanced by total ``weights`` (default: the
    capture lengths, i.e. samples).

    Rank r gets the captures whose cumulative-weight midpoint falls in the r-th 1/world_size
    slice of the total, so ranges are disjoint, ordered and cover [0, B)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    B = len(lengths)
    if world_size <= 1 or B == 0:
        return [(0, B)] + [(B, B)] * (max(world_size, 1) - 1)
    w = lengths.astype(np.float64) if weights is None else np.asarray(weights, dtype=np.float64)
    if len(w) != B:
        raise ValueError("weights must have one entry per capture")
    cum = np.cumsum(w)
    total = float(cum[-1])
    mids = cum - w / 2.0
    owner = np.minimum((mids * world_size / max(total, 1e-300)).astype(np.int64), world_size - 1)
    owner = np.maximum.accumulate(owner)
    bounds = np.searchsorted(owner, np.arange(world_size + 1), side="left")
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world_size)]


def imbalance(weights, ranges) -> float:
    """max / mean of the per-range weight sums (1.0 = perfectly balanced)."""
    w = np.asarray(weights, dtype=np.float64)
    sums = np.array([w[lo:hi].sum() for lo, hi in ranges])
    return float(sums.max() / max(sums.mean(), 1e-300))


def merge_rx_parts(parts):
    """[(results structured array, blob uint8, out_off int64[B_j+1]), ...] in corpus order -> the same
    triple for the whole corpus (payload offsets shifted by the blob bytes before each part)."""
    parts = list(parts)
    if not parts:
        raise ValueError("no parts")
    res = np.concatenate([p[0] for p in parts])
    blob_lo = np.cumsum([0] + [int(p[2][-1]) for p in parts]).astype(np.int64)
    blob = np.concatenate([np.asarray(p[1][:int(p[2][-1])], dtype=np.uint8) for p in parts]) if blob_lo[-1] else \
        np.zeros(0, np.uint8)
    out_off = np.concatenate([np.asarray(p[2][:-1], dtype=np.int64) + blob_lo[j] for j, p in enumerate(parts)] +
                             [blob_lo[-1:]]).astype(np.int64)
    return res, blob, out_off


def gather_batches(local, group=None):
    """All-gather arbitrary picklable per-rank results to every rank (torch.distributed, any backend).
    Returns the list ordered by rank; with no initialised process group returns [local]."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out


def gather_rx(batch, dst: int = 0, group=None):
    """Host gather of one RxBatch per rank (each the decode of that rank's contiguous shard, ranks in
    corpus order) into ONE RxBatch on rank ``dst`` (None elsewhere).  Three flat byte tensors per rank
    travel through torch.distributed (gloo or nccl): a few bytes per capture, off the data path."""
    import torch
    import torch.distributed as dist

    from .modem import RxBatch
    from ._cabi import RX_RESULT_DTYPE
    if not (dist.is_available() and dist.is_initialized()):
        return batch
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    used = int(batch.out_off[-1])
    res_b = np.ascontiguousarray(batch.results).view(np.uint8).reshape(-1)
    off_b = np.ascontiguousarray(batch.out_off, dtype=np.int64).view(np.uint8).reshape(-1)
    blob_b = np.ascontiguousarray(batch.blob[:used])
    nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    sizes = torch.tensor([len(res_b), len(off_b), len(blob_b)], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [[int(v) for v in s.cpu()] for s in all_sizes]
    flat = torch.from_numpy(np.concatenate([res_b, off_b, blob_b])).to(dev)
    # padded all_gather (every backend has it); the payload blobs are ~0.1 % of the samples
    width = max(sum(s) for s in all_sizes)
    pad = torch.zeros(width, dtype=torch.uint8, device=dev)
    pad[:flat.numel()] = flat
    bufs = [torch.zeros(width, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if rank != dst:
        return None
    parts = []
    for s, b in zip(all_sizes, bufs):
        h = b.cpu().numpy()
        r = h[:s[0]].view(RX_RESULT_DTYPE)
        o = h[s[0]:s[0] + s[1]].view(np.int64)
        parts.append((r, h[s[0] + s[1]:s[0] + s[1] + s[2]], o))
    return RxBatch(*merge_rx_parts(parts))
