"""Build recipe of libosb.so (the C-ABI CUDA library), in-tree, sm_100a only.

``python -m openset_imagenet_b200._build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels to the GPU box with the snapshot.
Every source is compiled to an object on its own (in parallel) and the objects are linked into one
shared library; a missing source is an error.  Build log and source fingerprint are written next to
the library (``libosb.build.log``, ``libosb.stamp``; both git-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libosb.so")
STAMP = os.path.join(CSRC, "libosb.stamp")
LOG = os.path.join(CSRC, "libosb.build.log")
OBJ_DIR = os.path.join(CSRC, "_obj")
SOURCES = ["lib.cu", "loss_rows.cu", "confidence.cu", "oscr.cu", "oscr_shard.cu", "hist.cu", "head_tc.cu", "dw_allreduce.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libosb.so cannot be built (there is no CPU fallback)")


def _fingerprint() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(" ".join(SOURCES).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                h.update(name.encode())
                h.update(open(os.path.join(root, name), "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu into one shared library; no-op when sources are unchanged."""
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == fp:
        return LIB
    missing = [s for s in SOURCES if not os.path.exists(os.path.join(CSRC, s))]
    if missing:
        raise RuntimeError(f"libosb.so: source file(s) missing from {CSRC}: {', '.join(missing)}")
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, cmd, proc

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = []
    failed = []
    for src, obj, cmd, proc in results:
        log.append(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if proc.returncode != 0:
            failed.append(src)
    if not failed:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *[r[1] for r in results]]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log.append(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if proc.returncode != 0:
            failed.append("link")
    text = "\n".join(log)
    with open(LOG, "w") as f:
        f.write(text)
    if failed:
        sys.stderr.write(text)
        raise RuntimeError(f"nvcc failed building libosb.so ({', '.join(failed)})")
    if verbose:
        print(text)
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
