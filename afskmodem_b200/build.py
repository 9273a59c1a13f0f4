"""Builds libafsk_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m afskmodem_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
working tree; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libafsk_b200.so")
SOURCES = ["afsk_api.cu", "afsk_rx.cu", "afsk_tx.cu", "afsk_io.cu"]
HEADERS = [os.path.join(CSRC, "afsk_common.cuh"), os.path.join(HERE, "..", "include", "afsk_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: tuple = ()) -> str:
    """out / defines: a tuning variant (e.g. defines=("AFSK_X=1",)) built beside the library of record."""
    if out is None and not force and not is_stale():
        return LIB
    target = out or LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
          ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return target


if __name__ == "__main__":
    # python -m afskmodem_b200.build [--force] [-v] [--out PATH] [-DNAME=VALUE ...]
    argv = sys.argv[1:]
    out = argv[argv.index("--out") + 1] if "--out" in argv else None
    print(build(force="--force" in argv, verbose="-v" in argv, out=out,
                defines=tuple(a[2:] for a in argv if a.startswith("-D"))))
