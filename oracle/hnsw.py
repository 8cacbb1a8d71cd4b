"""ctypes wrapper of oracle/hnsw_restatement.c -- BASELINE INFRASTRUCTURE ONLY.

An HNSW graph at the parameters pinned by the reference's shipped index
(vector_store/<segment>/header.bin: M = 16, maxM0 = 32, ef_construction = 100,
mult = 1/ln 16 = 0.360674) and Chroma's default search_ef = 10.  Used by
bench.py --hnsw-baseline to print "reference-style HNSW" recall and QPS, labelled as a
restatement, next to the exact search.  Never imported by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hnsw_restatement.c")
LIB = os.path.join(HERE, "_build", "libhnsw_restatement.so")

CHROMA_DEFAULTS = {"M": 16, "ef_construction": 100, "search_ef": 10, "space": "l2"}


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O3", "-march=x86-64-v3", "-ffast-math", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:       # retry without -march=native / openmp on exotic hosts
            cmd = ["gcc", "-O3", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("gcc failed for hnsw_restatement.c:\n" + r.stderr)
    return LIB


class HnswIndex:
    def __init__(self, data: np.ndarray, M: int = 16, ef_construction: int = 100, seed: int = 100):
        self._lib = C.CDLL(build())
        self._lib.hnsw_build.restype = C.c_void_p
        self._lib.hnsw_build.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64]
        self._lib.hnsw_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self._lib.hnsw_free.argtypes = [C.c_void_p]
        self._lib.hnsw_mult.restype = C.c_double
        self._lib.hnsw_mult.argtypes = [C.c_void_p]
        self._lib.hnsw_max_level.argtypes = [C.c_void_p]
        self.data = np.ascontiguousarray(data, dtype=np.float32)      # kept alive: the C side borrows it
        n, d = self.data.shape
        self._h = self._lib.hnsw_build(self.data.ctypes.data_as(C.c_void_p), n, d, M, ef_construction, seed)

    @property
    def mult(self) -> float:
        return float(self._lib.hnsw_mult(self._h))

    @property
    def max_level(self) -> int:
        return int(self._lib.hnsw_max_level(self._h))

    def query(self, queries: np.ndarray, k: int = 10, ef: int = 10):
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        ids = np.empty((q.shape[0], k), dtype=np.int64)
        d = np.empty((q.shape[0], k), dtype=np.float32)
        self._lib.hnsw_search(self._h, q.ctypes.data_as(C.c_void_p), q.shape[0], k, ef,
                              ids.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p))
        return ids, d

    def close(self):
        if self._h:
            self._lib.hnsw_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
