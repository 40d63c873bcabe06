"""Builds fesr_b200/lib/libfesr.so with nvcc for sm_100a (in-tree, so it travels with gpurun)."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfesr.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(src: str) -> str:
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    for f in sorted(os.listdir(CSRC)) + ["../../include/fesr.h"]:
        if f.endswith((".cuh", ".h")) or f == src:
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def _compile(src: str):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, "", False
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    with open(obj + ".ptxas.log", "w") as fh:
        fh.write(r.stderr)
    return obj, r.stderr, True


def build(verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    rebuilt = any(r[2] for r in results)
    if verbose:
        for (obj, log, did) in results:
            if did:
                print(f"[fesr build] compiled {os.path.basename(obj)}")
    if rebuilt or not os.path.exists(LIB):
        # -cudart shared: the runtime is the libcudart.so.12 the host process already has (PyTorch loads one), not a
        # private static copy inside libfesr.so; -ldl: comm.cu resolves NCCL with dlopen
        cmd = [NVCC, "-shared", "-cudart", "shared", "-o", LIB, *[r[0] for r in results],
               "-gencode", "arch=compute_100a,code=sm_100a", "-ldl",
               "-Xlinker", "-rpath", "-Xlinker", "/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[fesr build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
