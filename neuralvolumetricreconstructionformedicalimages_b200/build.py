"""Build libnafb200.so (hand-written sm_100a CUDA kernels behind the C ABI of include/nafb200.h).

In-tree build with explicit nvcc: the .so sits next to this file (git-ignored, but it travels
with the repository snapshot to the GPU box).  nvcc cross-compiles for sm_100a without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libnafb200.so")
DIAG_LIB = os.path.join(HERE, "libnafb200_diag.so")   # diagnostics (tcgen05 known-answer test, L2 micro-benchmarks): tests / scripts only
DIAG_SRC = os.path.join(CSRC, "diag")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=default",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for d in (CSRC, DIAG_SRC):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    for f in ("nafb200.h", "nafb200_diag.h"):
        with open(os.path.join(INCLUDE, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIB + ".stamp"
    if not (os.path.exists(LIB) and os.path.exists(DIAG_LIB) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libnafb200.so (+ csrc/diag/ -> libnafb200_diag.so). Returns its path."""
    if not force and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3].replace(os.sep, "_") + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    main_src = _sources()
    diag_src = sorted(os.path.join("diag", f) for f in os.listdir(DIAG_SRC) if f.endswith(".cu"))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, main_src + diag_src))
    for lib, lib_objs in ((LIB, objs[:len(main_src)]), (DIAG_LIB, objs[len(main_src):])):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *lib_objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(LIB + ".stamp", "w") as fh:
        fh.write(_fingerprint())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
