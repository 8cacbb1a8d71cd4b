"""Epilogue selection statistics of the tensor regime (RAG_B200_TENSOR_STATS=1): how often the warp-level fast
reject fails and how many list insertions a query batch costs.  Usage:
    RAG_B200_TENSOR_STATS=1 python tools/tensor_stats.py --rows 25000000 --dim 384 --space l2 --k 100 --batch 1024"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import local_rag_system_b200 as rag  # noqa: E402
from local_rag_system_b200 import _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=25_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--space", default="l2")
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--rerank", type=int, default=1)
ap.add_argument("--dtype", default="bf16")
a = ap.parse_args()
lib = _native.load()
st = rag.DeviceStore(a.dim, a.dtype, a.space, capacity_hint=a.rows, rerank=bool(a.rerank))
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
for s in range(0, a.rows, 500_000):
    m = min(500_000, a.rows - s)
    x = torch.nn.functional.normalize(torch.randn((m, a.dim), generator=gen, device="cuda"), dim=1)
    torch.cuda.synchronize()
    st.upsert_device(x.data_ptr(), m)
q = np.random.default_rng(0).standard_normal((a.batch, a.dim), dtype=np.float32)
q /= np.linalg.norm(q, axis=1, keepdims=True)
st.query(q, a.k)
out = (C.c_uint64 * 8)()
lib.rag_debug_tensor_stats(out, 1)
st.query(q, a.k)
ms = st.last_query_info()["kernel_ms"]
lib.rag_debug_tensor_stats(out, 1)
v = list(out)
tiles = max(v[0], 1)
print(f"rows {a.rows} dim {a.dim} {a.dtype} {a.space} k {a.k} B {a.batch} rerank {a.rerank}: kernel {ms:.3f} ms")
print(f"  warp-tiles drained {v[0]}, passed reject #1 {v[1]} ({100 * v[1] / tiles:.1f} %), passed #2 {v[2]} ({100 * v[2] / tiles:.1f} %)")
print(f"  candidate scores {v[3]} ({v[3] / tiles:.3f} per warp-tile), insertions {v[4]} ({v[4] / a.batch:.0f} per query), quantile updates {v[5]}")
