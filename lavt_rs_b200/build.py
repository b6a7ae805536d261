"""Build the C-ABI shared library (sm_100a only) in-tree with nvcc.

    python -m lavt_rs_b200.build            # incremental
    python -m lavt_rs_b200.build --force

The library has no PyTorch dependency; it is loaded with ctypes by ``lavt_rs_b200._cabi``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OUT_DIR = os.path.join(PKG_DIR, "_lib")
LIB_PATH = os.path.join(OUT_DIR, "liblavt_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("LAVT_NVCC_EXTRA", "").split()      # debug builds, e.g. LAVT_NVCC_EXTRA=-DT3_WATCHDOG


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return cand


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha1()
    # every header change rebuilds everything: cheap and safe
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(PKG_DIR, "..", "include", "lavt_b200.h"), "rb").read())
    h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in _sources():
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OUT_DIR, src[:-3] + ".o")
        stamp = obj + ".sha1"
        dig = _digest(sp)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp)
                and open(stamp).read() == dig):
            continue
        jobs.append((sp, obj, stamp, dig))

    def compile_one(job):
        sp, obj, stamp, dig = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", sp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = obj + ".log"
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {sp}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(dig)
        if verbose:
            print(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs,
               "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
