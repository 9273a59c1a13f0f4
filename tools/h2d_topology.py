#!/usr/bin/env python
"""Host -> device ingest ceiling of a multi-GPU box, and whether NVLink forwarding or NUMA placement lifts it.

    python tools/h2d_topology.py [--gib 2] [--chunk-mb 64]

One process drives every GPU with asynchronous copies on per-device streams (pinned host buffers,
cudaMemcpyAsync; peer copies over NVLink).  Prints one JSON line per experiment:

  all           every GPU copies its own buffer at once
  all_numa      same, every buffer allocated by a thread bound to the CPUs of the GPU's NUMA node
  even          only GPUs 0,2,4,.. copy (one per PCIe uplink pair)
  even_forward  GPUs 0,2,4,.. ingest their own buffer AND their odd sibling's, chunk by chunk, and forward the
                sibling's chunks with cudaMemcpyPeerAsync while the next chunk is on the PCIe link: every GPU
                ends up with its buffer, all host traffic goes through the even GPUs' links
  pairs_forward as even_forward for pairs (0,4),(1,5),.. — siblings on different uplinks (control)
"""
import argparse
import json
import os
import subprocess
import time

import torch


def gpu_numa_nodes(n):
    nodes = []
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout
        for line in out.strip().splitlines():
            idx, bus = [s.strip() for s in line.split(",")]
            bus = bus.lower()
            if bus.startswith("00000000:"):
                bus = "0000:" + bus[9:]
            p = f"/sys/bus/pci/devices/{bus}/numa_node"
            nodes.append(int(open(p).read()) if os.path.exists(p) else -1)
    except Exception:  # noqa: BLE001
        pass
    return (nodes + [-1] * n)[:n]


def node_cpus(node):
    p = f"/sys/devices/system/node/node{node}/cpulist"
    if node < 0 or not os.path.exists(p):
        return None
    cpus = set()
    for part in open(p).read().strip().split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def pinned(nbytes, cpus=None):
    old = os.sched_getaffinity(0)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus & old or old)
        except OSError:
            pass
    try:
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        t.fill_(1)                                   # touch every page from this thread
    finally:
        os.sched_setaffinity(0, old)
    return t


def run(name, devs, host, dev_buf, streams, chunk, forward=None, repeats=3):
    """devs: GPUs that copy from the host; forward: {ingesting GPU: sibling GPU} or None"""
    n = host[devs[0]].numel()
    best = 1e9
    for _ in range(repeats + 1):
        for d in range(len(dev_buf)):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for off in range(0, n, chunk):
            m = min(chunk, n - off)
            for d in devs:
                with torch.cuda.stream(streams[d][0]):
                    dev_buf[d][off:off + m].copy_(host[d][off:off + m], non_blocking=True)
                if forward:
                    s = forward[d]
                    slot = n + ((off // chunk) % 2) * chunk
                    stage = dev_buf[d][slot:slot + m]
                    with torch.cuda.stream(streams[d][0]):
                        # a staging half is free again once the forward copy that read it has finished (the forward of
                        # 64 MB over NVLink takes ~0.1 ms against 1.2 ms on the PCIe link)
                        streams[d][0].wait_stream(streams[d][1])
                        stage.copy_(host[s][off:off + m], non_blocking=True)
                    streams[d][1].wait_stream(streams[d][0])
                    with torch.cuda.stream(streams[d][1]):
                        dev_buf[s][off:off + m].copy_(stage, non_blocking=True)
        for d in range(len(dev_buf)):
            torch.cuda.synchronize(d)
        best = min(best, time.perf_counter() - t0)
    served = len(devs) * (2 if forward else 1)
    total = served * n / best / 1e9
    print(json.dumps({"experiment": name, "ingesting_gpus": devs, "gpus_served": served, "seconds": best,
                      "total_gbs": round(total, 1), "per_served_gpu_gbs": round(total / served, 1),
                      "per_ingesting_gpu_gbs": round(total / len(devs), 1)}), flush=True)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=2.0)
    ap.add_argument("--chunk-mb", type=int, default=64)
    args = ap.parse_args()
    ndev = torch.cuda.device_count()
    n = int(args.gib * (1 << 30))
    chunk = args.chunk_mb << 20
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
    nodes = gpu_numa_nodes(ndev)
    print(json.dumps({"gpus": ndev, "gpu_numa_node": nodes, "host_cpus": os.cpu_count(),
                      "numa_nodes": sorted({x for x in nodes if x >= 0})}), flush=True)
    dev_buf = [torch.empty(n + 2 * chunk, dtype=torch.uint8, device=f"cuda:{d}") for d in range(ndev)]
    streams = [[torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)] for d in range(ndev)]
    host = [pinned(n) for _ in range(ndev)]
    alld = list(range(ndev))
    run("one", [0], host, dev_buf, streams, chunk)
    run("all", alld, host, dev_buf, streams, chunk)
    if ndev >= 2:
        even = alld[0::2]
        run("even", even, host, dev_buf, streams, chunk)
        run("even_forward", even, host, dev_buf, streams, chunk, forward={d: d + 1 for d in even if d + 1 < ndev})
    if ndev >= 8:
        run("pairs_forward", alld[:4], host, dev_buf, streams, chunk, forward={d: d + 4 for d in alld[:4]})
        run("half_0123", alld[:4], host, dev_buf, streams, chunk)
    del host
    host = [pinned(n, node_cpus(nodes[d])) for d in range(ndev)]
    run("all_numa", alld, host, dev_buf, streams, chunk)
    if ndev >= 2:
        even = alld[0::2]
        run("even_numa", even, host, dev_buf, streams, chunk)
        run("even_forward_numa", even, host, dev_buf, streams, chunk, forward={d: d + 1 for d in even if d + 1 < ndev})


if __name__ == "__main__":
    main()
