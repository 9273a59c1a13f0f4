"""Admission control for host -> device copies on boxes where GPUs share host links.

Measured on an 8 x B200 box (tools/h2d_topology.py, profiles/r2_h2d_topology.txt): eight GPUs copying from
pinned host memory at once get 23.5 GB/s each (188 GB/s in total); four at once (one of every pair) get
54.7 GB/s each (219 GB/s); GPUs 0-3 alone share 115 GB/s.  The host side carries more when fewer links are
busy at a time, so the receiver lets only `slots` of every `group` consecutive GPUs copy at any moment: the
members of a group take turns range by range of a pipelined decode, each at the full link rate, while their
kernels overlap the other members' copies.  Measured end to end at 8 ranks (gpurun_out/r2d_n8, 4.9 GB per
rank): no gate 91.1 G samples/s (182 GB/s), 2:1 99.9 G (200 GB/s), 4:2 101.9 G (204 GB/s).  The gate is a set of lock files (flock), so it works between the threads of one process
(ShardedRxSession) and between one-process-per-GPU ranks alike; nothing is exchanged but the right to copy.

AFSK_H2D_GATE = "off" | "<group>:<slots>" overrides the default (4:2 from eight visible GPUs, off below).
"""
from __future__ import annotations

import fcntl
import os


def default_policy(visible_devices: int):
    """(group size, slots) or None"""
    ev = os.environ.get("AFSK_H2D_GATE", "").strip().lower()
    if ev in ("off", "0", "none"):
        return None
    if ":" in ev:
        g, s = ev.split(":", 1)
        return max(int(g), 1), max(int(s), 1)
    return (4, 2) if visible_devices >= 8 else None


class H2DGate:
    """``with gate:`` brackets one host -> device copy (including the wait for its completion)."""

    def __init__(self, device: int, group: int, slots: int, root: str | None = None):
        root = root or ("/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp")
        d = os.path.join(root, f"afsk_h2d_gate_{os.getuid()}")
        os.makedirs(d, exist_ok=True)
        self.group, self.group_id, self.member = group, device // group, device % group
        self.slots = min(slots, group)
        # one open file description per gate object: flock conflicts between descriptions, also inside a process
        self._fds = [os.open(os.path.join(d, f"g{self.group_id}_s{j}.lock"), os.O_RDWR | os.O_CREAT, 0o600)
                     for j in range(self.slots)]
        self._held = None

    def __enter__(self):
        for j, fd in enumerate(self._fds):                 # any free slot
            try:
                fcntl.flock(fd, fcntl.LOCK_EX | fcntl.LOCK_NB)
                self._held = j
                return self
            except OSError:
                continue
        j = self.member % self.slots                        # all busy: queue on this member's home slot
        fcntl.flock(self._fds[j], fcntl.LOCK_EX)
        self._held = j
        return self

    def __exit__(self, *exc):
        fcntl.flock(self._fds[self._held], fcntl.LOCK_UN)
        self._held = None
        return False

    def close(self):
        for fd in self._fds:
            try:
                os.close(fd)
            except OSError:
                pass
        self._fds = []

    __del__ = close


def make_gate(device: int):
    """The gate a session on `device` should use on this machine, or None."""
    from . import _cabi
    pol = default_policy(_cabi.device_count())
    return H2DGate(device, *pol) if pol else None
