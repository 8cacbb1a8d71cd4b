"""Stream vs tensor regime for small batches on one store (sets tensor::kStreamMaxBatch).
usage: crossover.py ROWS DIM DTYPE"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import local_rag_system_b200 as rag  # noqa: E402
from local_rag_system_b200.sharded import ShardedSearcher  # noqa: E402

rows, dim, dtype = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
dev = torch.device("cuda", 0)
st = rag.DeviceStore(dim, dtype, "cosine", capacity_hint=rows)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
for s in range(0, rows, 500_000):
    m = min(500_000, rows - s)
    x = torch.randn((m, dim), generator=gen, device=dev)
    torch.cuda.synchronize()
    st.upsert_device(x.data_ptr(), m)
se = ShardedSearcher(st, 0, 1)
for B in (1, 2, 3, 4, 6, 8, 12, 16):
    q = torch.randn((40, B, dim), generator=gen, device=dev)
    line = [f"B={B:2d}"]
    for regime in ("stream", "tensor"):
        try:
            for i in range(5):
                se.search_device(q[i], 10, regime=regime)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(5, 40):
                se.search_device(q[i], 10, regime=regime)
            e1.record()
            torch.cuda.synchronize()
            line.append(f"{regime} {e0.elapsed_time(e1) / 35:.4f} ms")
        except Exception as e:  # noqa: BLE001
            line.append(f"{regime} n/a ({type(e).__name__})")
    print("  ".join(line), flush=True)
st.close()
