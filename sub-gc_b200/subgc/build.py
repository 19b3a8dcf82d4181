"""Build libsubgc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m subgc.build [--force]

The library sits next to this file so that it travels with the repository snapshot to the GPU box; it is
git-ignored, and so is the stamp beside it (a content hash of the sources + flags over repo-relative paths): a checkout
never carries a stamp that vouches for somebody else's binary.  The build is skipped when library and stamp match the
sources; `subgc._lib.lib()` additionally refuses a library whose ABI version differs from the binding's.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
INCLUDE = os.path.join(ROOT, "include")
OUT = os.path.join(HERE, "libsubgc_b200.so")
STAMP = OUT + ".stamp"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]
FLAGS += os.environ.get("SUBGC_NVCC_EXTRA", "").split()   # experiments only (e.g. -DMG_LEAN); part of the stamp


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256(" ".join(FLAGS).encode())
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files += [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())   # repo-relative: the digest does not depend on where the checkout lives
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.isfile(OUT) and os.path.isfile(STAMP) and open(STAMP).read().strip() == digest:
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        # box without a toolkit: only a library whose stamp matches these sources may be used (it travelled with the snapshot)
        if os.path.isfile(OUT) and os.path.isfile(STAMP) and open(STAMP).read().strip() == digest:
            return OUT
        raise RuntimeError("nvcc not found and no libsubgc_b200.so built from these sources is present")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-I", CSRC, "-o", OUT] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libsubgc_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
