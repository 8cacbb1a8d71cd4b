"""Small driver for ncu captures of the tensor-regime kernel: builds a synthetic store and runs a few
B = 1024 batches.  usage: prof_tensor.py ROWS DIM [SPACE] [K]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import local_rag_system_b200 as rag  # noqa: E402

rows, dim = int(sys.argv[1]), int(sys.argv[2])
space = sys.argv[3] if len(sys.argv) > 3 else "cosine"
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dtype = sys.argv[5] if len(sys.argv) > 5 else "bf16"
B = int(sys.argv[6]) if len(sys.argv) > 6 else 1024
dev = torch.device("cuda", 0)
st = rag.DeviceStore(dim, dtype, space, capacity_hint=rows)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
for s in range(0, rows, 500_000):
    m = min(500_000, rows - s)
    x = torch.randn((m, dim), generator=gen, device=dev)
    torch.cuda.synchronize()
    st.upsert_device(x.data_ptr(), m)
q = np.random.default_rng(2).standard_normal((B, dim)).astype(np.float32)
for _ in range(4):
    r, d, c = st.query(q, k)
print("ok", st.last_query_info(), r[0, :3])
st.close()
