"""Stress the fp32 tensor regime: many batches on one store, every launch synchronous (run with
CUDA_LAUNCH_BLOCKING=1 to pin a failing launch to the call the engine reports).
    python tools/stress_f32.py [rows] [dim] [shadow] [iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import local_rag_system_b200 as rag  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
shadow = sys.argv[3] if len(sys.argv) > 3 else "hi"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 200
rng = np.random.default_rng(0)
dtype = os.environ.get("STRESS_DTYPE", "f32")
st = rag.DeviceStore(dim, dtype, "cosine", capacity_hint=rows)
for s0 in range(0, rows, 250_000):
    st.upsert(rng.standard_normal((min(250_000, rows - s0), dim), dtype=np.float32))
if dtype == "f32":
    st.set_f32_shadow(shadow)
for B in [int(v) for v in os.environ.get("STRESS_B", "32,1024,200").split(",")]:
    q = rng.standard_normal((4, B, dim), dtype=np.float32)
    t0 = time.time()
    for i in range(iters):
        try:
            r, d, c = st.query(q[i % 4], 10, regime="tensor")
        except Exception as e:  # noqa: BLE001
            print(f"B={B} iter {i}: {e}", flush=True)
            raise
        if i == 0:
            first = r.copy()
        elif i % 4 == 0:
            assert np.array_equal(first, r), f"B={B} iter {i}: result changed between identical batches"
    print(f"B={B}: {iters} batches ok, {1e3 * (time.time() - t0) / iters:.3f} ms per blocking call, info {st.f32_tensor_info() if dtype == "f32" else st.last_query_info()}", flush=True)
st.close()
