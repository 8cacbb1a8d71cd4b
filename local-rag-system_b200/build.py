"""Build librag_b200.so (the C-ABI engine) in-tree with nvcc for sm_100a.

    python local-rag-system_b200/build.py [--force]

Objects go to local-rag-system_b200/build/, the library next to this file.
nvcc cross-compiles without a GPU, so this runs in the CPU-only build
container; the resulting .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "librag_b200.so")
SOURCES = ["api.cu", "sharded.cu", "scan_stream.cu", "merge.cu", "store_kernels.cu", "tensor_regime.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the engine cannot be built")


def _deps_mtime() -> float:
    t = os.path.getmtime(os.path.join(HERE, "..", "include", "rag_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".h", ".cuh")):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def _compile(nvcc: str, src: str, force: bool, hdr_t: float) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    srcp = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), hdr_t)):
        return obj
    cmd = [nvcc, *NVCC_FLAGS, "-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _deps_mtime()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s, force, hdr_t), SOURCES))
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
               "-Xlinker", "--no-undefined", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
